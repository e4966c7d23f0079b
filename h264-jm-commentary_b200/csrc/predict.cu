// predict.cu — MV prediction (H.264 8.4.1.3) and the ME-only commit of a motion field (sm_100a).
//
// Stands in for JM's SetMotionVectorPredictor ‖ GetMotionVectorPredictorNormal
// evaluated for all 41 blocks of every macroblock against a committed 4x4-granular field
// (JM enc_picture->mv / ref_idx), and for the closed ME-only mode decision of SURVEY.md §7 H1 that
// produces such a field from a search result.  Negligible arithmetic; the kernels exist so that the
// predictor -> search -> commit loop of consecutive frames never leaves the device-side library.
#include <cuda_runtime.h>

#include <cstdlib>

#include "jmme_dev.cuh"

namespace {

struct Nb {          // one neighbour: vector, reference index (-1 = none), availability
    int x, y, ref, avail;
};

__host__ __device__ inline int med3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

// 8.4.1.3: directional rules of 16x8 / 8x16, else 8.4.1.3.1 median rules.  C is already D when C is missing.
__host__ __device__ inline void mv_predict(int t, int part, int ref, Nb A, Nb B, Nb C, int &px, int &py)
{
    if (!A.avail || A.ref < 0) { A.x = A.y = 0; A.ref = -1; }
    if (!B.avail || B.ref < 0) { B.x = B.y = 0; B.ref = -1; }
    if (!C.avail || C.ref < 0) { C.x = C.y = 0; C.ref = -1; }
    const Nb *dir = nullptr;
    if (t == 2) dir = part == 0 ? &B : &A;
    if (t == 3) dir = part == 0 ? &A : &C;
    if (dir && dir->ref == ref) { px = dir->x; py = dir->y; return; }
    if (!B.avail && !C.avail && A.avail) { B = A; C = A; }
    const int hit = (A.ref == ref) + (B.ref == ref) + (C.ref == ref);
    if (hit == 1) {
        const Nb &m = A.ref == ref ? A : (B.ref == ref ? B : C);
        px = m.x; py = m.y;
    } else {
        px = med3(A.x, B.x, C.x); py = med3(A.y, B.y, C.y);
    }
}

// decoded-before test of field cell (x,y) for the block (t, x0, y0) of MB (mbx, mby)
__device__ inline int cell_ready(int x, int y, int fw, int fh, int mbx, int mby, int t, int x0, int y0)
{
    if (x < 0 || y < 0 || x >= fw || y >= fh) return 0;
    const int mx = x >> 2, my = y >> 2;
    if (my != mby) return my < mby;
    if (mx != mbx) return mx < mbx;
    const int lx = (x & 3) * 4, ly = (y & 3) * 4;
    if (t == 1) return 0;
    if (t == 2) return (ly >> 3) < (y0 >> 3);
    if (t == 3) return (lx >> 3) < (x0 >> 3);
    const int bw = t <= 5 ? 8 : 4, bh = (t == 4 || t == 6) ? 8 : 4;
    const int q = 2 * (y0 >> 3) + (x0 >> 3), qc = 2 * (ly >> 3) + (lx >> 3);
    if (qc != q) return qc < q;
    const int s = ((y0 & 7) / bh) * (8 / bw) + (x0 & 7) / bw, sc = ((ly & 7) / bh) * (8 / bw) + (lx & 7) / bw;
    return sc < s;
}

__global__ void predict_kernel(const int16_t *__restrict__ mv4, const int8_t *__restrict__ ref4, int mb_w, int mb_h,
                               int num_refs, int16_t *__restrict__ pred)
{
    const int n_mb = mb_w * mb_h;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_refs * n_mb * JMME_NBLK) return;
    const int blk = i % JMME_NBLK, mb = (i / JMME_NBLK) % n_mb, ref = i / (JMME_NBLK * n_mb);
    const int mbx = mb % mb_w, mby = mb / mb_w, fw = 4 * mb_w, fh = 4 * mb_h;
    const int t = c_blk_type[blk], x0 = c_blk_x[blk], y0 = c_blk_y[blk];
    const int cx = 4 * mbx + (x0 >> 2), cy = 4 * mby + (y0 >> 2), wc = c_blk_w[blk] >> 2;
    const int nx[4] = {cx - 1, cx, cx + wc, cx - 1}, ny[4] = {cy, cy - 1, cy - 1, cy - 1};
    Nb nb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        nb[k].avail = cell_ready(nx[k], ny[k], fw, fh, mbx, mby, t, x0, y0);
        nb[k].x = nb[k].y = 0; nb[k].ref = -1;
        if (nb[k].avail) {
            const size_t o = (size_t)ny[k] * fw + nx[k];
            nb[k].x = mv4[2 * o]; nb[k].y = mv4[2 * o + 1]; nb[k].ref = ref4[o];
        }
    }
    if (!nb[2].avail) nb[2] = nb[3];
    int px, py;
    mv_predict(t, t == 2 ? (y0 >> 3) : (t == 3 ? (x0 >> 3) : 0), ref, nb[0], nb[1], nb[2], px, py);
    pred[2 * (size_t)i] = (int16_t)px;
    pred[2 * (size_t)i + 1] = (int16_t)py;
}

// one thread per MB: cheapest of 16x16 / 16x8 / 8x16 / P8x8 (sub-type per 8x8), then the 16 field cells
__global__ void commit_kernel(const jmme_mbresult *__restrict__ res, int mb_w, int mb_h, int mask,
                              int16_t *__restrict__ mv4, int8_t *__restrict__ ref4, uint8_t *__restrict__ mode)
{
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= mb_w * mb_h) return;
    const jmme_mbresult &m = res[mb];
    const long long INF = 0x7FFFFFFFFFFFFFFFll;
    long long J[4];
    int sub[4] = {0, 0, 0, 0};
    J[0] = (mask >> 1) & 1 ? (long long)m.cost[0] : INF;
    J[1] = (mask >> 2) & 1 ? (long long)m.cost[1] + m.cost[2] : INF;
    J[2] = (mask >> 3) & 1 ? (long long)m.cost[3] + m.cost[4] : INF;
    J[3] = 0;
    for (int q = 0; q < 4; q++) {
        long long bq = INF;
        for (int t = 4; t <= 7; t++) {
            if (!((mask >> t) & 1)) continue;
            long long s = 0;
            for (int b = 5; b < JMME_NBLK; b++)          // blocks of type t inside 8x8 number q
                if (c_blk_type[b] == t && 2 * (c_blk_y[b] >> 3) + (c_blk_x[b] >> 3) == q) s += m.cost[b];
            if (s < bq) { bq = s; sub[q] = t; }
        }
        if (bq == INF) { J[3] = INF; break; }
        J[3] += bq;
    }
    int md = 0;
    for (int k = 1; k < 4; k++) if (J[k] < J[md]) md = k;
    mode[5 * mb] = (uint8_t)(md == 3 ? 8 : md + 1);
    for (int q = 0; q < 4; q++) mode[5 * mb + 1 + q] = (uint8_t)(md == 3 ? sub[q] : 0);
    const int mbx = mb % mb_w, mby = mb / mb_w, fw = 4 * mb_w;
    for (int cell = 0; cell < 16; cell++) {
        const int cx4 = cell & 3, cy4 = cell >> 2;
        const int t = md == 3 ? sub[2 * (cy4 >> 1) + (cx4 >> 1)] : md + 1;
        int blk = 0;
        for (int b = 0; b < JMME_NBLK; b++)              // the block of type t that covers this cell
            if (c_blk_type[b] == t && 4 * cx4 >= c_blk_x[b] && 4 * cx4 < c_blk_x[b] + c_blk_w[b] &&
                4 * cy4 >= c_blk_y[b] && 4 * cy4 < c_blk_y[b] + c_blk_h[b])
                blk = b;
        const size_t o = (size_t)(4 * mby + cy4) * fw + 4 * mbx + cx4;
        mv4[2 * o] = m.mv[blk][0]; mv4[2 * o + 1] = m.mv[blk][1]; ref4[o] = m.ref_idx[blk];
    }
}

}  // namespace

cudaError_t jmme_launch_predict(const int16_t *mv4, const int8_t *ref4, int mb_w, int mb_h, int num_refs, int16_t *pred,
                                cudaStream_t st)
{
    const int n = num_refs * mb_w * mb_h * JMME_NBLK;
    predict_kernel<<<(n + 255) / 256, 256, 0, st>>>(mv4, ref4, mb_w, mb_h, num_refs, pred);
    return cudaGetLastError();
}

cudaError_t jmme_launch_commit(const jmme_mbresult *res, int mb_w, int mb_h, int mask, int16_t *mv4, int8_t *ref4,
                               uint8_t *mode, cudaStream_t st)
{
    const int n = mb_w * mb_h;
    commit_kernel<<<(n + 127) / 128, 128, 0, st>>>(res, mb_w, mb_h, mask, mv4, ref4, mode);
    return cudaGetLastError();
}

extern "C" int jmme_SetMotionVectorPredictor(int blocktype, int part, int ref_idx, const int16_t mvA[2], int refA,
                                             int availA, const int16_t mvB[2], int refB, int availB,
                                             const int16_t mvC[2], int refC, int availC, int16_t pred[2])
{
    if (blocktype < 1 || blocktype > 7 || !mvA || !mvB || !mvC || !pred) return JMME_ERR_PARAM;
    Nb A{mvA[0], mvA[1], refA, availA}, B{mvB[0], mvB[1], refB, availB}, C{mvC[0], mvC[1], refC, availC};
    int px, py;
    mv_predict(blocktype, part, ref_idx, A, B, C, px, py);    // scalar host evaluation of the device routine
    pred[0] = (int16_t)px; pred[1] = (int16_t)py;
    return JMME_OK;
}
