// me_bipred.cu — bi-predictive refinement of the vector pairs of two uni-directional searches (sm_100a).
//
// Stands in for JM's BiPredBlockMotionSearch / the *BiPred* search of me_fullsearch.c (SURVEY.md §8(f) rank 2;
// names recalled, nothing of it exists under /root/reference).  Definition: include/jmme.h,
// jmme_search_frame_bipred; DESIGN.md §2 "bi-predictive refinement".
//
// Mapping: one CTA (4 warps) per macroblock; the 41 blocks one after the other (every block has its own pair of
// vectors, so nothing is shared between them); per iteration the fixed list's prediction block is staged in
// shared memory, thread = candidate of the searched list: reference words through L1/L2 (unaligned: two loads
// and a funnel shift), prediction = per-byte rounded average (__vavgu4 = (a + b + 1) >> 1, the standard's default
// weighted prediction), distortion VABSDIFF4.ACC (or VABSDIFF4 + IDP.4A for SSE), 64-bit (cost, position) minimum
// over the CTA.  Position 0 (the current pair) is candidate 0, so the spiral order with strict < is the order
// of the packed keys.
#include "jmme_dev.cuh"

struct BipredArgs {
    const uint8_t *planes_l1;            // the list-1 picture: n_planes padded planes
    const jmme_mbresult *l0, *l1;        // uni-directional results, whole-frame indexed
    const int16_t *pred0, *pred1;        // predictors of the two lists (context's policy layout) or null
    const int16_t *spiral_xy;            // [(2 range + 1)^2][2]
    int range, iterations, npb, n_planes;
    jmme_bipred *out;
    int *err;                            // set to 1 when a record cannot be used (reference index, vector phase)
};

namespace {

__global__ void __launch_bounds__(128) me_bipred_kernel(const SearchParams P, const BipredArgs A)
{
    __shared__ uint32_t s_cur[64], s_fix[64];
    __shared__ unsigned long long s_red[4];
    __shared__ int s_mv[2][2], s_cost;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mb = d_unit_mb(P, blockIdx.x);
    const int mby = mb / P.mb_w, mbx = mb - mby * P.mb_w;
    const int n_mb = P.mb_w * P.mb_h;
    const int ncand = (2 * A.range + 1) * (2 * A.range + 1);
    const size_t psz = (size_t)P.pstride * P.pheight;
    const int dom = P.cost_domain, lf0 = P.lf[0];
    const bool sse = P.metric[0] == JMME_DIST_SSE;
    jmme_bipred *o = A.out + mb;

    if (tid < 64) {
        const int row = tid >> 2, w = tid & 3;
        s_cur[tid] = *(const uint32_t *)(P.cur + (size_t)min(16 * mby + row, P.cur_h - 1) * P.cur_stride + 16 * mbx + 4 * w);
    }
    if (tid < 3) o->reserved[tid] = 0;
    __syncthreads();

    for (int b = 0; b < JMME_NBLK; b++) {
        const bool on = (P.blocktype_mask >> c_blk_type[b]) & 1;
        const int r0 = on ? A.l0[mb].ref_idx[b] : -1;
        const bool usable = on && r0 >= 0 && r0 < P.num_refs && A.l1[mb].ref_idx[b] >= 0;
        if (!usable) {                                   // (uniform over the CTA)
            if (tid == 0) {
                o->mv0[b][0] = o->mv0[b][1] = o->mv1[b][0] = o->mv1[b][1] = 0;
                o->cost[b] = INT_MAX; o->ref0[b] = -1;
                if (on) *A.err = 1;
            }
            continue;
        }
        const int bx = c_blk_x[b], by = c_blk_y[b], bw4 = c_blk_w[b] >> 2, bh = c_blk_h[b];
        int pr[2][2] = {{0, 0}, {0, 0}};
        if (A.pred0) {
            const int16_t *q = A.pred0 + ((size_t)r0 * n_mb + mb) * A.npb * 2 + (A.npb == 1 ? 0 : 2 * b);
            pr[0][0] = d_pred(q[0]); pr[0][1] = d_pred(q[1]);
        }
        if (A.pred1) {
            const int16_t *q = A.pred1 + (size_t)mb * A.npb * 2 + (A.npb == 1 ? 0 : 2 * b);
            pr[1][0] = d_pred(q[0]); pr[1][1] = d_pred(q[1]);
        }
        if (tid == 0) {
            s_mv[0][0] = A.l0[mb].mv[b][0]; s_mv[0][1] = A.l0[mb].mv[b][1];
            s_mv[1][0] = A.l1[mb].mv[b][0]; s_mv[1][1] = A.l1[mb].mv[b][1];
            s_cost = INT_MAX;
            if (A.n_planes == 1 && ((s_mv[0][0] | s_mv[0][1] | s_mv[1][0] | s_mv[1][1]) & 3)) *A.err = 1;
        }
        __syncthreads();
        const uint8_t *pl[2] = {P.planes[r0], A.planes_l1};
        // block position at vector (0,0) in padded-plane coordinates
        const int rx0 = P.pad + 16 * mbx + bx, ry0 = P.pad + 16 * mby + by;
        for (int it = 0; it < A.iterations; it++) {
            const int s = it & 1, f = 1 - s;
            const int fx = s_mv[f][0], fy = s_mv[f][1], sx0 = s_mv[s][0], sy0 = s_mv[s][1];
            // sub-pel vectors need the 16 planes; with the integer plane alone the phase bits are ignored (and flagged above)
            const int fph = A.n_planes == 1 ? 0 : (fy & 3) * 4 + (fx & 3), sph = A.n_planes == 1 ? 0 : (sy0 & 3) * 4 + (sx0 & 3);
            if (tid < bh * bw4) {                        // the fixed list's prediction block
                const int y = tid / bw4, w = tid - y * bw4;
                const size_t off = psz * fph + (size_t)(ry0 + (fy >> 2) + y) * P.pstride + (rx0 + (fx >> 2) + 4 * w);
                const uint32_t *rp = (const uint32_t *)(pl[f] + (off & ~(size_t)3));
                s_fix[4 * y + w] = __funnelshift_r(__ldg(rp), __ldg(rp + 1), (int)(off & 3) * 8);
            }
            __syncthreads();
            const int fbits = d_se_bits(fx - pr[f][0]) + d_se_bits(fy - pr[f][1]);
            unsigned long long best = ~0ull;
            for (int pos = tid; pos < ncand; pos += 128) {
                const int qx = sx0 + 4 * A.spiral_xy[2 * pos], qy = sy0 + 4 * A.spiral_xy[2 * pos + 1];
                if ((qx >> 2) < -(P.pad - 1) || (qx >> 2) > P.pad - 1 || (qy >> 2) < -(P.pad - 1) || (qy >> 2) > P.pad - 1) continue;
                const size_t off = psz * sph + (size_t)(ry0 + (qy >> 2)) * P.pstride + (rx0 + (qx >> 2));
                const uint32_t *rp = (const uint32_t *)(pl[s] + (off & ~(size_t)3));
                const int sh = (int)(off & 3) * 8, pw = P.pstride >> 2;
                unsigned d = 0;
                for (int y = 0; y < bh; y++) {
                    uint32_t a = __ldg(rp + y * pw);
                    for (int w = 0; w < bw4; w++) {
                        const uint32_t nx = __ldg(rp + y * pw + w + 1);
                        const uint32_t p = __vavgu4(s_fix[4 * y + w], __funnelshift_r(a, nx, sh));
                        const uint32_t cw = s_cur[4 * (by + y) + (bx >> 2) + w];
                        if (sse) { const unsigned e = __vabsdiffu4(cw, p); d = __dp4a(e, e, d); }
                        else d = sad4(cw, p, d);
                        a = nx;
                    }
                }
                const int c = d_dscale(dom, (int)d) + d_wcost(dom, lf0, fbits + d_se_bits(qx - pr[s][0]) + d_se_bits(qy - pr[s][1]));
                const unsigned long long v = ((unsigned long long)(unsigned)(c + 0x40000000) << 32) | (unsigned)pos;
                best = v < best ? v : best;
            }
            for (int sft = 16; sft; sft >>= 1) {
                const unsigned long long u = __shfl_xor_sync(0xFFFFFFFFu, best, sft);
                best = u < best ? u : best;
            }
            if (lane == 0) s_red[warp] = best;
            __syncthreads();
            if (tid == 0) {
                unsigned long long m = s_red[0];
                for (int k = 1; k < 4; k++) m = s_red[k] < m ? s_red[k] : m;
                if (m != ~0ull) {
                    const int pos = (int)(unsigned)m;
                    s_mv[s][0] = sx0 + 4 * A.spiral_xy[2 * pos]; s_mv[s][1] = sy0 + 4 * A.spiral_xy[2 * pos + 1];
                    s_cost = (int)(unsigned)(m >> 32) - 0x40000000;
                }
            }
            __syncthreads();
        }
        if (tid == 0) {
            o->mv0[b][0] = (int16_t)s_mv[0][0]; o->mv0[b][1] = (int16_t)s_mv[0][1];
            o->mv1[b][0] = (int16_t)s_mv[1][0]; o->mv1[b][1] = (int16_t)s_mv[1][1];
            o->cost[b] = s_cost; o->ref0[b] = (int8_t)r0;
        }
        __syncthreads();
    }
}

}  // namespace

cudaError_t jmme_launch_bipred(const SearchParams &P, const BipredArgs &A, cudaStream_t st)
{
    me_bipred_kernel<<<d_n_units(P), 128, 0, st>>>(P, A);
    return cudaGetLastError();
}
