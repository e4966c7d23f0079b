// me_full.cu — FullPelBlockMotionSearch with a window of its own per block (sm_100a).
//
// Stands in for JM's FullPelBlockMotionSearch in the one configuration the
// shared-window kernels (me_int*.cu) cannot express: search_mode = FULL with 41 different predictors
// per macroblock, where every block's window is centred on its own predictor.  No SAD can be shared
// between blocks then, so the work is 7 x 256 abs-diffs per candidate and MB instead of 256; this
// kernel is the plain, correct form (thread = candidate, VABSDIFF4 on unaligned words fetched through
// L1), kept off the headline path.  Conventions: DESIGN.md §2 (spiral order, strict <, 16x16 bonus;
// no (0,0) pre-test in FULL mode).
#include <cstdio>

#include "jmme_dev.cuh"

namespace {

__global__ void __launch_bounds__(128) me_full_kernel(const SearchParams P)
{
    __shared__ uint32_t s_cur[64];
    __shared__ unsigned long long s_best[JMME_NBLK];

    const int tid = threadIdx.x, lane = tid & 31;
    const int R = P.R, ncols = P.ncols, ncand = ncols * ncols;
    const int n_mb_stripe = d_n_units(P);
    const int n_mb = P.mb_w * P.mb_h;
    const int item = blockIdx.x;
    const int ref = item / n_mb_stripe;
    const int mb = d_unit_mb(P, item - ref * n_mb_stripe);
    const int mby = mb / P.mb_w, mbx = mb - mby * P.mb_w;
    const int npb = P.pred_policy == JMME_PRED_PER_BLOCK ? JMME_NBLK : 1;
    const int16_t *pr = P.pred ? P.pred + ((size_t)ref * n_mb + mb) * npb * 2 : nullptr;
    const int dom = P.cost_domain, lf0 = P.lf[0];
    const bool sse = P.metric[0] == JMME_DIST_SSE;
    const int bonus16 = (!P.rdopt && ref == 0) ? d_wcost(dom, lf0, 16) : 0;
    const uint8_t *plane = P.planes[ref];

    if (tid < 64) {
        const int row = tid >> 2, w = tid & 3;
        s_cur[tid] = *(const uint32_t *)(P.cur + (size_t)min(16 * mby + row, P.cur_h - 1) * P.cur_stride + 16 * mbx + 4 * w);
    }
    if (tid < JMME_NBLK) s_best[tid] = ~0ull;
    __syncthreads();

    for (int b = 0; b < JMME_NBLK; b++) {
        if (!((P.blocktype_mask >> c_blk_type[b]) & 1)) continue;
        const int bx = c_blk_x[b], by = c_blk_y[b], bw4 = c_blk_w[b] >> 2, bh = c_blk_h[b];
        const int px = pr ? d_pred(pr[2 * (npb == 1 ? 0 : b)]) : 0, py = pr ? d_pred(pr[2 * (npb == 1 ? 0 : b) + 1]) : 0;
        const int cx = d_clamp(px / 4, -P.cmax, P.cmax), cy = d_clamp(py / 4, -P.cmax, P.cmax);
        const int bonus = b == 0 ? bonus16 : 0;
        unsigned long long best = ~0ull;
        for (int idx = tid; idx < ncand; idx += 128) {
            const int yoff = idx / ncols, xoff = idx - yoff * ncols;
            const int mx = cx + xoff - R, my = cy + yoff - R;
            const size_t off = (size_t)(P.pad + 16 * mby + by + my) * P.pstride + (P.pad + 16 * mbx + bx + mx);
            const uint32_t *rp = (const uint32_t *)(plane + (off & ~(size_t)3));
            const int sh = (int)(off & 3) * 8, pw = P.pstride >> 2;
            unsigned s = 0;
            for (int y = 0; y < bh; y++) {
                uint32_t a = __ldg(rp + y * pw);
                for (int w = 0; w < bw4; w++) {
                    const uint32_t nx = __ldg(rp + y * pw + w + 1);
                    const uint32_t cw = s_cur[4 * (by + y) + (bx >> 2) + w], rw = __funnelshift_r(a, nx, sh);
                    if (sse) { const unsigned d = __vabsdiffu4(cw, rw); s = __dp4a(d, d, s); }
                    else s = sad4(cw, rw, s);
                    a = nx;
                }
            }
            int c = d_dscale(dom, (int)s) + d_wcost(dom, lf0, d_se_bits(4 * mx - px) + d_se_bits(4 * my - py));
            if (mx == 0 && my == 0) c -= bonus;
            const unsigned long long v = ((unsigned long long)(unsigned)(c + 0x40000000) << 32) | P.spiral_key[idx];
            best = v < best ? v : best;
        }
        for (int o = 16; o; o >>= 1) {
            const unsigned long long u = __shfl_xor_sync(0xFFFFFFFFu, best, o);
            best = u < best ? u : best;
        }
        if (lane == 0) atomicMin(&s_best[b], best);
    }
    __syncthreads();
    if (tid < JMME_NBLK && ((P.blocktype_mask >> c_blk_type[tid]) & 1)) {
        const unsigned long long v = s_best[tid];
        const unsigned key = (unsigned)v;
        const int px = pr ? d_pred(pr[2 * (npb == 1 ? 0 : tid)]) : 0, py = pr ? d_pred(pr[2 * (npb == 1 ? 0 : tid) + 1]) : 0;
        const int cx = d_clamp(px / 4, -P.cmax, P.cmax), cy = d_clamp(py / 4, -P.cmax, P.cmax);
        BlkRes r;
        r.mvx = (int16_t)(4 * (cx + P.spiral_xy[2 * (key - 1)]));
        r.mvy = (int16_t)(4 * (cy + P.spiral_xy[2 * (key - 1) + 1]));
        r.cost = (int)(unsigned)(v >> 32) - 0x40000000;
        P.res[((size_t)ref * n_mb + mb) * JMME_NBLK + tid] = r;
    }
}

}  // namespace

cudaError_t jmme_launch_me_full(const SearchParams &P, cudaStream_t st)
{
    const int n_items = d_n_units(P) * P.num_refs;
    snprintf(jmme_kernel_name_buf(), JMME_KNAME_LEN, "me_full_kernel");
    me_full_kernel<<<n_items, 128, 0, st>>>(P);
    return cudaGetLastError();
}
