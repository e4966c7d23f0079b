// leaf.cu — the JM-named leaf entry points of include/jmme.h on plain host arrays.
//
// Function-by-function parity surface: each call copies its small inputs to the
// device, runs a dedicated kernel and copies the answer back.  These are not the throughput path
// (that is jmme_search_frame); they exist so that every stage can be checked against the oracle
// in isolation and so that a JM-side shim has a one-to-one target for each function it replaces.
#include <cuda_runtime.h>

#include <cstdlib>
#include <vector>

#include "jmme_dev.cuh"

cudaError_t jmme_launch_interp(const uint8_t *src, int w_in, int h_in, int stride, int pad, int ps, int ph,
                               int n_planes, uint8_t *out, int y_begin, int y_end, cudaStream_t st);

namespace {

struct DevBuf {
    void *p = nullptr;
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
    ~DevBuf() { cudaFree(p); }
    template <class T> T *as() { return (T *)p; }
};
#define LCU(call)                                         \
    do {                                                  \
        cudaError_t e_ = (call);                          \
        if (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) return JMME_ERR_NODEVICE; \
        if (e_ != cudaSuccess) return JMME_ERR_CUDA;      \
    } while (0)

__device__ __forceinline__ int satd_of(const int *d, int satd_round)
{
    int t[16], s = 0;
    for (int i = 0; i < 4; i++) {
        int a = d[4 * i], b = d[4 * i + 1], c = d[4 * i + 2], e = d[4 * i + 3];
        int s0 = a + e, s1 = b + c, d0 = a - e, d1 = b - c;
        t[4 * i] = s0 + s1; t[4 * i + 1] = d0 + d1; t[4 * i + 2] = s0 - s1; t[4 * i + 3] = d0 - d1;
    }
    for (int i = 0; i < 4; i++) {
        int a = t[i], b = t[4 + i], c = t[8 + i], e = t[12 + i];
        int s0 = a + e, s1 = b + c, d0 = a - e, d1 = b - c;
        s += abs(s0 + s1) + abs(d0 + d1) + abs(s0 - s1) + abs(d0 - d1);
    }
    return satd_round ? (s + 1) >> 1 : s >> 1;
}

__global__ void satd_kernel(const int16_t *diff, int n, int satd_round, int32_t *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d[16];
    for (int k = 0; k < 16; k++) d[k] = diff[16 * (size_t)i + k];
    out[i] = satd_of(d, satd_round);
}

// JM HadamardSAD8x8: one thread per 8x8 difference block, three butterfly stages per row then per column
__global__ void satd8_kernel(const int16_t *diff, int n, int satd_round, int32_t *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int m[64];
    for (int k = 0; k < 64; k++) m[k] = diff[64 * (size_t)i + k];
    for (int pass = 0; pass < 2; pass++) {
        const int sr = pass ? 1 : 8, se = pass ? 8 : 1;          // stride between lines / between elements of a line
        for (int l = 0; l < 8; l++) {
            int v[8];
            for (int k = 0; k < 8; k++) v[k] = m[l * sr + k * se];
            for (int h = 4; h; h >>= 1)
                for (int k = 0; k < 8; k++)
                    if (!(k & h)) { const int a = v[k], b = v[k + h]; v[k] = a + b; v[k + h] = a - b; }
            for (int k = 0; k < 8; k++) m[l * sr + k * se] = v[k];
        }
    }
    int s = 0;
    for (int k = 0; k < 64; k++) s += abs(m[k]);
    out[i] = satd_round ? (s + 2) >> 2 : s >> 2;
}

// the 64 eighth-pel planes of one chroma component [STD 8.4.2.2.2]; thread = padded sample, loops over the phases
__global__ void chroma_planes_kernel(const uint8_t *src, int w, int h, int stride, int pad, int ps, int ph, uint8_t *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= ps || y >= ph) return;
    const int x0 = d_clamp(x - pad, 0, w - 1), x1 = d_clamp(x + 1 - pad, 0, w - 1);
    const int y0 = d_clamp(y - pad, 0, h - 1), y1 = d_clamp(y + 1 - pad, 0, h - 1);
    const int A = src[(size_t)y0 * stride + x0], B = src[(size_t)y0 * stride + x1];
    const int C = src[(size_t)y1 * stride + x0], D = src[(size_t)y1 * stride + x1];
    const size_t psz = (size_t)ps * ph;
    for (int yf = 0; yf < 8; yf++)
        for (int xf = 0; xf < 8; xf++)
            out[psz * (yf * 8 + xf) + (size_t)y * ps + x] =
                (uint8_t)(((8 - xf) * (8 - yf) * A + xf * (8 - yf) * B + (8 - xf) * yf * C + xf * yf * D + 32) >> 6);
}

// one thread per spiral position: 16 4x4 SADs with VABSDIFF4 on byte-gathered words, then the sums
__global__ void blocksad_kernel(const uint8_t *cur, const uint8_t *ref, int rs, int ox, int oy, int cx, int cy,
                                int ncand, const int16_t *sxy, int bonus, int32_t *out)
{
    int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= ncand) return;
    const int mx = cx + sxy[2 * pos], my = cy + sxy[2 * pos + 1];
    unsigned s[16];
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++) {
            unsigned acc = 0;
            for (int y = 0; y < 4; y++) {
                const uint8_t *c = cur + (4 * j + y) * 16 + 4 * i;
                const uint8_t *r = ref + (size_t)(oy + my + 4 * j + y) * rs + (ox + mx + 4 * i);
                unsigned cw = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
                unsigned rw = r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);
                acc = sad4(cw, rw, acc);
            }
            s[4 * j + i] = acc;
        }
#define O(b) out[(size_t)(b) * ncand + pos]
    for (int i = 0; i < 16; i++) O(25 + i) = s[i];
    for (int j = 0; j < 2; j++) for (int i = 0; i < 4; i++) O(17 + 4 * j + i) = s[8 * j + i] + s[8 * j + 4 + i];
    for (int j = 0; j < 4; j++) for (int i = 0; i < 2; i++) O(9 + 2 * j + i) = s[4 * j + 2 * i] + s[4 * j + 2 * i + 1];
    unsigned q[4];
    for (int j = 0; j < 2; j++)
        for (int i = 0; i < 2; i++) {
            q[2 * j + i] = s[8 * j + 2 * i] + s[8 * j + 2 * i + 1] + s[8 * j + 4 + 2 * i] + s[8 * j + 4 + 2 * i + 1];
            O(5 + 2 * j + i) = q[2 * j + i];
        }
    O(3) = q[0] + q[2]; O(4) = q[1] + q[3];
    O(1) = q[0] + q[1]; O(2) = q[2] + q[3];
    O(0) = (int)(q[0] + q[1] + q[2] + q[3]) - ((mx == 0 && my == 0) ? bonus : 0);
#undef O
}

// single-CTA argmin over a cost surface with the MV rate; 64-bit (cost, key) keys keep JM's order
__device__ __forceinline__ unsigned long long pack64(int cost, unsigned key)
{
    return ((unsigned long long)(unsigned)(cost + 0x40000000) << 32) | key;
}
__device__ unsigned long long block_min64(unsigned long long v, unsigned long long *sh)
{
    for (int o = 16; o; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, v, o);
        v = w < v ? w : v;
    }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : ~0ull;
        for (int o = 16; o; o >>= 1) {
            unsigned long long w = __shfl_xor_sync(0xFFFFFFFFu, v, o);
            v = w < v ? w : v;
        }
        if (threadIdx.x == 0) sh[0] = v;
    }
    __syncthreads();
    return sh[0];
}

// mode 0: costs come from blocksad[] (a7); mode 1: SAD of a bw x bh block computed here (a8)
__global__ void __launch_bounds__(256) fullpel_kernel(int mode, const int32_t *blocksad, const uint8_t *cur, int cs,
                                                      const uint8_t *ref, int rs, int bx, int by, int bw, int bh, int cx,
                                                      int cy, int px, int py, int ncand, const int16_t *sxy, int f,
                                                      int bonus, int pretest, int *out3)
{
    __shared__ unsigned long long sh[32];
    unsigned long long best = ~0ull;
    for (int pos = threadIdx.x; pos < ncand; pos += blockDim.x) {
        const int mx = cx + sxy[2 * pos], my = cy + sxy[2 * pos + 1];
        int c = d_weighted_cost(f, d_se_bits(4 * mx - px) + d_se_bits(4 * my - py));
        if (mode == 0) {
            c += blocksad[pos];
        } else {
            int s = 0;
            for (int y = 0; y < bh; y++)
                for (int x = 0; x < bw; x++)
                    s += abs((int)cur[(size_t)(by + y) * cs + bx + x] - (int)ref[(size_t)(by + my + y) * rs + bx + mx + x]);
            c += s;
            if (mx == 0 && my == 0) c -= bonus;
        }
        unsigned key = (pretest && mx == 0 && my == 0) ? 0u : (unsigned)pos + 1u;
        unsigned long long v = pack64(c, key);
        best = v < best ? v : best;
    }
    best = block_min64(best, sh);
    if (threadIdx.x == 0) {
        unsigned key = (unsigned)best;
        int pos = -1;
        if (key) pos = (int)key - 1;
        out3[0] = pos < 0 ? 0 : cx + sxy[2 * pos];
        out3[1] = pos < 0 ? 0 : cy + sxy[2 * pos + 1];
        out3[2] = (int)(unsigned)(best >> 32) - 0x40000000;
    }
}

__constant__ int8_t c_leaf_sp9[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};

// one block, one CTA: thread = (position, 4x4 cell)
__global__ void __launch_bounds__(256) subpel_leaf_kernel(const uint8_t *cur, int cs, const uint8_t *planes, int ps,
                                                          int ph, int pad, int bx, int by, int bw, int bh, int px, int py,
                                                          int f, int hadamard, int satd_round, int bonus, int *io3)
{
    __shared__ int s_cost[9];
    __shared__ int s_mv[2], s_min;
    const int ncx = bw / 4, ncells = ncx * (bh / 4);
    if (threadIdx.x == 0) { s_mv[0] = io3[0]; s_mv[1] = io3[1]; s_min = hadamard ? INT_MAX : io3[2]; }
    for (int step = 2; step >= 1; step--) {
        const int pos0 = (step == 2 && hadamard) ? 0 : 1;
        if (threadIdx.x < 9) s_cost[threadIdx.x] = 0;
        __syncthreads();
        for (int u = threadIdx.x; u < 9 * ncells; u += blockDim.x) {
            const int pos = u / ncells, cell = u - pos * ncells;
            if (pos < pos0) continue;
            const int x4 = 4 * (cell % ncx), y4 = 4 * (cell / ncx);
            const int qx = s_mv[0] + step * c_leaf_sp9[pos][0], qy = s_mv[1] + step * c_leaf_sp9[pos][1];
            const uint8_t *rp = planes + (size_t)ps * ph * ((qy & 3) * 4 + (qx & 3)) +
                                (size_t)(pad + by + y4 + (qy >> 2)) * ps + (pad + bx + x4 + (qx >> 2));
            int d[16], v = 0;
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++)
                    d[4 * y + x] = (int)cur[(size_t)(by + y4 + y) * cs + bx + x4 + x] - (int)rp[(size_t)y * ps + x];
            if (hadamard) v = satd_of(d, satd_round);
            else for (int k = 0; k < 16; k++) v += abs(d[k]);
            atomicAdd(&s_cost[pos], v);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int mn = s_min, best = 0;
            const int ox = s_mv[0], oy = s_mv[1];
            for (int pos = pos0; pos < 9; pos++) {
                const int qx = ox + step * c_leaf_sp9[pos][0], qy = oy + step * c_leaf_sp9[pos][1];
                int c = d_weighted_cost(f, d_se_bits(qx - px) + d_se_bits(qy - py)) + s_cost[pos];
                if (qx == 0 && qy == 0) c -= bonus;
                if (c < mn) { mn = c; best = pos; }
            }
            s_min = mn; s_mv[0] = ox + step * c_leaf_sp9[best][0]; s_mv[1] = oy + step * c_leaf_sp9[best][1];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { io3[0] = s_mv[0]; io3[1] = s_mv[1]; io3[2] = s_min; }
}

int spiral_index(int dx, int dy)
{
    if (!dx && !dy) return 0;
    int l = std::max(std::abs(dx), std::abs(dy)), base = (2 * l - 1) * (2 * l - 1);
    if (std::abs(dy) == l && std::abs(dx) < l) return base + 2 * (dx + l - 1) + (dy > 0);
    return base + 2 * (2 * l - 1) + 2 * (dy + l) + (dx > 0);
}
void spiral_table(int R, std::vector<int16_t> &xy)
{
    const int n = (2 * R + 1) * (2 * R + 1);
    xy.assign(2 * (size_t)n, 0);
    for (int dy = -R; dy <= R; dy++)
        for (int dx = -R; dx <= R; dx++) {
            int k = spiral_index(dx, dy);
            xy[2 * (size_t)k] = (int16_t)dx; xy[2 * (size_t)k + 1] = (int16_t)dy;
        }
}
int se_bits_host(int v)
{
    int a = std::abs(v), k = 0;
    if (!a) return 1;
    while ((2 << k) <= a) k++;
    return 2 * k + 3;
}

// copy a w x h byte rectangle (host, strided) into a dense device buffer
cudaError_t upload_rect(DevBuf &d, const uint8_t *src, int stride, int w, int h)
{
    cudaError_t e = d.alloc((size_t)w * h);
    if (e != cudaSuccess) return e;
    return cudaMemcpy2D(d.p, w, src, stride, w, h, cudaMemcpyHostToDevice);
}

}  // namespace

extern "C" {

int jmme_InitMotionSearchModule(int R, int max_mvd, int32_t *mvbits, int n_refbits, int32_t *refbits, int16_t *sx,
                                int16_t *sy)
{
    if (R < 0 || max_mvd < 0 || n_refbits < 0) return JMME_ERR_PARAM;
    if (mvbits) for (int v = -max_mvd; v <= max_mvd; v++) mvbits[v + max_mvd] = se_bits_host(v);
    if (refbits)
        for (int r = 0; r < n_refbits; r++) {
            int k = 0;
            while ((2 << k) <= r + 1) k++;
            refbits[r] = 2 * k + 1;
        }
    if (sx && sy) {
        std::vector<int16_t> xy;
        spiral_table(R, xy);
        for (size_t i = 0; i < xy.size() / 2; i++) { sx[i] = xy[2 * i]; sy[i] = xy[2 * i + 1]; }
    }
    return JMME_OK;
}

int jmme_getSubImagesLuma(const uint8_t *luma, int width, int height, int stride, int pad, uint8_t *out_planes)
{
    if (!luma || !out_planes || width <= 0 || height <= 0 || (width & 15) || (height & 15) || pad < 0 ||
        stride < width)
        return JMME_ERR_PARAM;
    if (pad & 3) return JMME_ERR_UNSUPPORTED;           // planes are written as 32-bit words
    const int ps = width + 2 * pad, ph = height + 2 * pad;
    DevBuf src, dst;
    LCU(upload_rect(src, luma, stride, width, height));
    LCU(dst.alloc((size_t)ps * ph * 16));
    LCU(jmme_launch_interp(src.as<uint8_t>(), width, height, width, pad, ps, ph, 16, dst.as<uint8_t>(), 0, ph, 0));
    LCU(cudaMemcpy(out_planes, dst.p, (size_t)ps * ph * 16, cudaMemcpyDeviceToHost));
    return JMME_OK;
}

int jmme_HadamardSAD8x8(const int16_t *diff, int n, int satd_round, int32_t *out)
{
    if (!diff || !out || n < 0) return JMME_ERR_PARAM;
    if (n == 0) return JMME_OK;
    DevBuf d_in, d_out;
    LCU(d_in.alloc(sizeof(int16_t) * 64 * (size_t)n));
    LCU(d_out.alloc(sizeof(int32_t) * (size_t)n));
    LCU(cudaMemcpy(d_in.p, diff, sizeof(int16_t) * 64 * (size_t)n, cudaMemcpyHostToDevice));
    satd8_kernel<<<(n + 63) / 64, 64>>>(d_in.as<int16_t>(), n, satd_round, d_out.as<int32_t>());
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(out, d_out.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    return JMME_OK;
}

int jmme_getSubImagesChroma(const uint8_t *chroma, int width, int height, int stride, int pad, uint8_t *out)
{
    if (!chroma || !out || width <= 0 || height <= 0 || pad < 0 || stride < width) return JMME_ERR_PARAM;
    const int ps = width + 2 * pad, ph = height + 2 * pad;
    DevBuf d_in, d_out;
    LCU(d_in.alloc((size_t)stride * height));
    LCU(d_out.alloc((size_t)ps * ph * 64));
    LCU(cudaMemcpy(d_in.p, chroma, (size_t)stride * (height - 1) + width, cudaMemcpyHostToDevice));
    chroma_planes_kernel<<<dim3((ps + 127) / 128, ph), 128>>>(d_in.as<uint8_t>(), width, height, stride, pad, ps, ph,
                                                             d_out.as<uint8_t>());
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(out, d_out.p, (size_t)ps * ph * 64, cudaMemcpyDeviceToHost));
    return JMME_OK;
}

int jmme_SATD(const int16_t *diff, int n, int satd_round, int32_t *out)
{
    if (!diff || !out || n < 0) return JMME_ERR_PARAM;
    if (n == 0) return JMME_OK;
    DevBuf d, o;
    LCU(d.alloc(sizeof(int16_t) * 16 * (size_t)n));
    LCU(o.alloc(sizeof(int32_t) * (size_t)n));
    LCU(cudaMemcpy(d.p, diff, sizeof(int16_t) * 16 * (size_t)n, cudaMemcpyHostToDevice));
    satd_kernel<<<(n + 255) / 256, 256>>>(d.as<int16_t>(), n, satd_round, o.as<int32_t>());
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(out, o.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    return JMME_OK;
}

int jmme_SetupFastFullPelSearch(const uint8_t *cur, int cs, const uint8_t *ref, int rs, int mbx, int mby, int cx, int cy,
                                int R, int bonus, int32_t *out)
{
    if (!cur || !ref || !out || R < 1 || R > JMME_MAX_SEARCH_RANGE) return JMME_ERR_PARAM;
    if (abs(cx) > R || abs(cy) > R) return JMME_ERR_PARAM;
    const int n = (2 * R + 1) * (2 * R + 1), win = 4 * R + 16;
    // window of every reachable sample: rows/cols [-2R, 2R+16) around the MB origin
    const uint8_t *w0 = ref + (ptrdiff_t)(16 * mby - 2 * R) * rs + (16 * mbx - 2 * R);
    DevBuf dc, dr, dxy, dout;
    std::vector<int16_t> xy;
    spiral_table(R, xy);
    LCU(upload_rect(dc, cur, cs, 16, 16));
    LCU(upload_rect(dr, w0, rs, win, win));
    LCU(dxy.alloc(xy.size() * 2));
    LCU(dout.alloc(sizeof(int32_t) * JMME_NBLK * (size_t)n));
    LCU(cudaMemcpy(dxy.p, xy.data(), xy.size() * 2, cudaMemcpyHostToDevice));
    blocksad_kernel<<<(n + 127) / 128, 128>>>(dc.as<uint8_t>(), dr.as<uint8_t>(), win, 2 * R, 2 * R, cx, cy, n,
                                              dxy.as<int16_t>(), bonus, dout.as<int32_t>());
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(out, dout.p, sizeof(int32_t) * JMME_NBLK * (size_t)n, cudaMemcpyDeviceToHost));
    return JMME_OK;
}

int jmme_FastFullPelBlockMotionSearch(const int32_t *sad, int R, int cx, int cy, int px, int py, int f, int pretest,
                                      int16_t *mvx, int16_t *mvy, int32_t *cost)
{
    if (!sad || !mvx || !mvy || !cost || R < 1 || R > JMME_MAX_SEARCH_RANGE) return JMME_ERR_PARAM;
    if (abs(px) > JMME_MAX_PRED_QPEL || abs(py) > JMME_MAX_PRED_QPEL) return JMME_ERR_PARAM;
    const int n = (2 * R + 1) * (2 * R + 1);
    DevBuf ds, dxy, dout;
    std::vector<int16_t> xy;
    int o3[3];
    spiral_table(R, xy);
    LCU(ds.alloc(sizeof(int32_t) * (size_t)n));
    LCU(dxy.alloc(xy.size() * 2));
    LCU(dout.alloc(sizeof o3));
    LCU(cudaMemcpy(ds.p, sad, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(dxy.p, xy.data(), xy.size() * 2, cudaMemcpyHostToDevice));
    fullpel_kernel<<<1, 256>>>(0, ds.as<int32_t>(), nullptr, 0, nullptr, 0, 0, 0, 0, 0, cx, cy, px, py, n,
                               dxy.as<int16_t>(), f, 0, pretest, dout.as<int>());
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(o3, dout.p, sizeof o3, cudaMemcpyDeviceToHost));
    *mvx = (int16_t)o3[0]; *mvy = (int16_t)o3[1]; *cost = o3[2];
    return JMME_OK;
}

int jmme_FullPelBlockMotionSearch(const uint8_t *cur, int cs, const uint8_t *ref, int rs, int bx, int by, int bw, int bh,
                                  int px, int py, int R, int f, int bonus, int16_t *mvx, int16_t *mvy, int32_t *cost)
{
    if (!cur || !ref || !mvx || !mvy || !cost || R < 1 || R > JMME_MAX_SEARCH_RANGE || bw <= 0 || bh <= 0 || bw > 16 ||
        bh > 16)
        return JMME_ERR_PARAM;
    if (abs(px) > JMME_MAX_PRED_QPEL || abs(py) > JMME_MAX_PRED_QPEL) return JMME_ERR_PARAM;
    const int n = (2 * R + 1) * (2 * R + 1), wx = 4 * R + bw, wy = 4 * R + bh;
    const int cx = std::min(std::max(px / 4, -R), R), cy = std::min(std::max(py / 4, -R), R);
    const uint8_t *w0 = ref + (ptrdiff_t)(by - 2 * R) * rs + (bx - 2 * R);
    DevBuf dc, dr, dxy, dout;
    std::vector<int16_t> xy;
    int o3[3];
    spiral_table(R, xy);
    LCU(upload_rect(dc, cur + (size_t)by * cs + bx, cs, bw, bh));
    LCU(upload_rect(dr, w0, rs, wx, wy));
    LCU(dxy.alloc(xy.size() * 2));
    LCU(dout.alloc(sizeof o3));
    LCU(cudaMemcpy(dxy.p, xy.data(), xy.size() * 2, cudaMemcpyHostToDevice));
    // device buffers are re-based so that block position (0,0) is at offset 2R inside the window
    fullpel_kernel<<<1, 256>>>(1, nullptr, dc.as<uint8_t>(), bw, dr.as<uint8_t>() + (size_t)2 * R * wx + 2 * R, wx, 0, 0,
                               bw, bh, cx, cy, px, py, n, dxy.as<int16_t>(), f, bonus, 0, dout.as<int>());
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(o3, dout.p, sizeof o3, cudaMemcpyDeviceToHost));
    *mvx = (int16_t)o3[0]; *mvy = (int16_t)o3[1]; *cost = o3[2];
    return JMME_OK;
}

int jmme_SubPelBlockMotionSearch(const uint8_t *cur, int cs, const uint8_t *planes, int width, int height, int pad, int bx,
                                 int by, int bw, int bh, int px, int py, int f, int hadamard, int satd_round, int bonus,
                                 int16_t *mvx, int16_t *mvy, int32_t *cost)
{
    if (!cur || !planes || !mvx || !mvy || !cost || bw <= 0 || bh <= 0 || (bw & 3) || (bh & 3)) return JMME_ERR_PARAM;
    if (abs(px) > JMME_MAX_PRED_QPEL || abs(py) > JMME_MAX_PRED_QPEL) return JMME_ERR_PARAM;
    const int ps = width + 2 * pad, ph = height + 2 * pad;
    DevBuf dc, dp, dio;
    int io3[3] = {*mvx, *mvy, *cost};
    LCU(upload_rect(dc, cur + (size_t)by * cs + bx, cs, bw, bh));
    LCU(dp.alloc((size_t)ps * ph * 16));
    LCU(dio.alloc(sizeof io3));
    LCU(cudaMemcpy(dp.p, planes, (size_t)ps * ph * 16, cudaMemcpyHostToDevice));
    LCU(cudaMemcpy(dio.p, io3, sizeof io3, cudaMemcpyHostToDevice));
    // the current block was uploaded densely: address it as block (0,0) with stride bw, but keep the
    // reference addressing at (bx,by)
    subpel_leaf_kernel<<<1, 256>>>(dc.as<uint8_t>() - ((size_t)by * bw + bx), bw, dp.as<uint8_t>(), ps, ph, pad, bx, by, bw,
                                   bh, px, py, f, hadamard, satd_round, bonus, dio.as<int>());
    LCU(cudaGetLastError());
    LCU(cudaMemcpy(io3, dio.p, sizeof io3, cudaMemcpyDeviceToHost));
    *mvx = (int16_t)io3[0]; *mvy = (int16_t)io3[1]; *cost = io3[2];
    return JMME_OK;
}

}  // extern "C"
