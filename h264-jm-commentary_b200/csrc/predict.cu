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
#include "wave.cuh"

namespace {

// decoded-before test of field cell (x,y) for the block (t, x0, y0) of MB (mbx, mby)
__host__ __device__ inline int cell_ready(int x, int y, int fw, int fh, int mbx, int mby, int t, int x0, int y0)
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

// predictor of block `blk` of MB (mbx,mby) for reference `ref` from the field (mv4, ref4).
// slice_top: first MB row of the slice; p16 non-null: the in-frame median policy, where the
// already-decoded partitions of this MB itself carry (p16, ref).
__device__ inline void predict_one(const int16_t *mv4, const int8_t *ref4, int mb_w, int mb_h, int mbx, int mby, int blk,
                                   int ref, int slice_top, const int16_t *p16, int &px, int &py)
{
    const int fw = 4 * mb_w, fh = 4 * mb_h;
    const int t = c_blk_type[blk], x0 = c_blk_x[blk], y0 = c_blk_y[blk];
    const int cx = 4 * mbx + (x0 >> 2), cy = 4 * mby + (y0 >> 2), wc = c_blk_w[blk] >> 2;
    const int nx[4] = {cx - 1, cx, cx + wc, cx - 1}, ny[4] = {cy, cy - 1, cy - 1, cy - 1};
    Nb nb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        nb[k].avail = cell_ready(nx[k], ny[k], fw, fh, mbx, mby, t, x0, y0) && ny[k] >= 4 * slice_top;
        nb[k].x = nb[k].y = 0; nb[k].ref = -1;
        if (nb[k].avail) {
            if (p16 && (nx[k] >> 2) == mbx && (ny[k] >> 2) == mby) {
                nb[k].x = p16[0]; nb[k].y = p16[1]; nb[k].ref = ref;
            } else {
                const size_t o = (size_t)ny[k] * fw + nx[k];
                nb[k].x = mv4[2 * o]; nb[k].y = mv4[2 * o + 1]; nb[k].ref = ref4[o];
            }
        }
    }
    if (!nb[2].avail) nb[2] = nb[3];
    mv_predict(t, t == 2 ? (y0 >> 3) : (t == 3 ? (x0 >> 3) : 0), ref, nb[0], nb[1], nb[2], px, py);
}

__global__ void predict_kernel(const int16_t *__restrict__ mv4, const int8_t *__restrict__ ref4, int mb_w, int mb_h,
                               int num_refs, int16_t *__restrict__ pred)
{
    const int n_mb = mb_w * mb_h;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_refs * n_mb * JMME_NBLK) return;
    const int blk = i % JMME_NBLK, mb = (i / JMME_NBLK) % n_mb, ref = i / (JMME_NBLK * n_mb);
    int px, py;
    predict_one(mv4, ref4, mb_w, mb_h, mb % mb_w, mb / mb_w, blk, ref, 0, nullptr, px, py);
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
    const int mo = mb_mode(m.cost, mask), md = mo & 3;
    mode[5 * mb] = (uint8_t)(md == 3 ? 8 : md + 1);
    for (int q = 0; q < 4; q++) mode[5 * mb + 1 + q] = (uint8_t)(md == 3 ? (mo >> (4 + 4 * q)) & 15 : 0);
    const int mbx = mb % mb_w, mby = mb / mb_w, fw = 4 * mb_w;
    for (int cell = 0; cell < 16; cell++) {
        const int cx4 = cell & 3, cy4 = cell >> 2;
        const int blk = cell_block(mo, cx4, cy4);
        const size_t o = (size_t)(4 * mby + cy4) * fw + 4 * mbx + cx4;
        mv4[2 * o] = m.mv[blk][0]; mv4[2 * o + 1] = m.mv[blk][1]; ref4[o] = m.ref_idx[blk];
    }
}

// In-frame median (JMME_PRED_MEDIAN), the predictors of one step of the 2:1 wavefront, a single CTA.  (The
// MBs of the previous step were committed to the field by the kernels that wrote their records: me_subpel.cu.)
// The step is on the serial path of the frame (predict -> search -> commit, step after step), so every phase
// fetches what it needs with one round of independent loads into shared memory and computes from there:
//   1. the 10 neighbour cells (4 left, 6 above) of every MB of this step,
//   2. the 16x16 predictor of every (reference, MB) of this step,
//   3. the other 40 predictors, whose in-MB neighbours carry that 16x16 predictor.
#define WAVE_CH 128                                       // MBs per chunk
__global__ void __launch_bounds__(1024) wave_step_kernel(const int *__restrict__ cur, int n_cur, int mb_w, int mb_h,
                                                         int num_refs, int slice_rows, const int16_t *mv4,
                                                         const int8_t *ref4, int16_t *pred, const WaveTab tab)
{
    __shared__ WaveNb s_nb[WAVE_CH][10];
    __shared__ int16_t s_p16[JMME_MAX_REFS][WAVE_CH][2];
    __shared__ int s_mb[WAVE_CH];
    __shared__ uint8_t s_src[JMME_NBLK][4], s_tp[JMME_NBLK];
    const int tid = threadIdx.x, n_mb = mb_w * mb_h, fw = 4 * mb_w, fh = 4 * mb_h;
    if (tid < JMME_NBLK * 4) s_src[tid >> 2][tid & 3] = tab.src[tid >> 2][tid & 3];
    if (tid < JMME_NBLK) s_tp[tid] = tab.tp[tid];
    const int k = slice_rows ? slice_rows : mb_h;
    for (int c0 = 0; c0 < n_cur; c0 += WAVE_CH) {
        const int n = min(WAVE_CH, n_cur - c0);
        if (tid < n) s_mb[tid] = cur[c0 + tid];
        __syncthreads();
        for (int i = tid; i < 10 * n; i += 1024) {         // slot 0..3: left column, 4..9: the row above from x-1
            const int j = i / 10, sl = i - 10 * j, mb = s_mb[j], mby = mb / mb_w, mbx = mb - mby * mb_w;
            int x, y;
            wave_slot_xy(mbx, mby, sl, x, y);
            s_nb[j][sl] = wave_load_nb(mv4, ref4, fw, fh, 4 * ((mby / k) * k), x, y);
        }
        __syncthreads();
        auto predict = [&](int j, int blk, int ref, int p16x, int p16y, int &px, int &py) {
            wave_predict_block(s_nb[j], s_src[blk], s_tp[blk], ref, p16x, p16y, px, py);
        };
        for (int i = tid; i < n * num_refs; i += 1024) {
            const int ref = i / n, j = i - ref * n;
            int px, py;
            predict(j, 0, ref, 0, 0, px, py);
            s_p16[ref][j][0] = (int16_t)px; s_p16[ref][j][1] = (int16_t)py;
        }
        __syncthreads();
        for (int i = tid; i < n * num_refs * JMME_NBLK; i += 1024) {
            const int blk = i % JMME_NBLK, q = i / JMME_NBLK;
            const int ref = q / n, j = q - ref * n, mb = s_mb[j];
            int px = s_p16[ref][j][0], py = s_p16[ref][j][1];
            if (blk) predict(j, blk, ref, px, py, px, py);
            *(uint32_t *)(pred + (((size_t)ref * n_mb + mb) * JMME_NBLK + blk) * 2) = (uint32_t)(uint16_t)px | ((uint32_t)(uint16_t)py << 16);
        }
        __syncthreads();
    }
}

// the neighbour-source table of wave_step_kernel, from the same decoding-order rule as predict_kernel
WaveTab build_wave_tab()
{
    WaveTab T;
    static const int bw_[8] = {0, 16, 16, 8, 8, 8, 4, 4}, bh_[8] = {0, 16, 8, 16, 8, 4, 8, 4}, base_[8] = {0, 0, 1, 3, 5, 9, 17, 25};
    const int mbx = 1, mby = 1, X = 4, Y = 4;               // an MB with every outer neighbour inside a 3x3-MB picture
    for (int t = 1; t <= 7; t++)
        for (int j = 0; j < 16 / bh_[t]; j++)
            for (int i = 0; i < 16 / bw_[t]; i++) {
                const int blk = base_[t] + j * (16 / bw_[t]) + i, x0 = i * bw_[t], y0 = j * bh_[t];
                const int cx = X + x0 / 4, cy = Y + y0 / 4, wc = bw_[t] / 4;
                const int nx[4] = {cx - 1, cx, cx + wc, cx - 1}, ny[4] = {cy, cy - 1, cy - 1, cy - 1};
                for (int q = 0; q < 4; q++) {
                    const bool inside = nx[q] >= X && nx[q] < X + 4 && ny[q] >= Y;
                    const int ready = cell_ready(nx[q], ny[q], 12, 12, mbx, mby, t, x0, y0);
                    int src;
                    if (inside) src = ready ? 10 : 11;
                    else if (!ready) src = 11;                // the MB to the right: decoded later
                    else src = ny[q] >= Y ? ny[q] - Y : 4 + nx[q] - (X - 1);
                    T.src[blk][q] = (uint8_t)src;
                }
                T.tp[blk] = (uint8_t)(t | ((t == 2 ? j : (t == 3 ? i : 0)) << 3));
            }
    return T;
}
}  // namespace
const WaveTab &jmme_wave_tab()
{
    static const WaveTab T = build_wave_tab();                // thread-safe one-time initialisation
    return T;
}
namespace {

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

cudaError_t jmme_launch_wave_step(const int *cur, int n_cur, int mb_w, int mb_h, int num_refs, int slice_rows,
                                  const int16_t *mv4, const int8_t *ref4, int16_t *pred, cudaStream_t st)
{
    wave_step_kernel<<<1, 1024, 0, st>>>(cur, n_cur, mb_w, mb_h, num_refs, slice_rows, mv4, ref4, pred, jmme_wave_tab());
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
