// me_full.cu — FullPelBlockMotionSearch with a window of its own per block (sm_100a).
//
// Stands in for JM's FullPelBlockMotionSearch in the one configuration the
// shared-window kernels (me_int*.cu) cannot express: search_mode = FULL with 41 different predictors
// per macroblock, where every block's window is centred on its own predictor.  No SAD can be shared
// between blocks then: 7 x 256 abs-diffs per candidate and MB instead of 256.  Conventions: DESIGN.md §2
// (spiral order, strict <, 16x16 bonus; no (0,0) pre-test in FULL mode).
//
// One CTA (4 warps) per (reference, MB); the 41 blocks one after the other:
//   window   the MB's (16 + 2R)^2 reference bytes around the staged centre come in as aligned 16-byte chunks and are
//            expanded into one 32-bit word per byte position (word[row][x] = bytes x..x+3, the layout of
//            me_int_tb.cu): a candidate at column x reads words x, x+4, ... — aligned, no shift in front of
//            VABSDIFF4, consecutive lanes on consecutive banks.  A block reads the window at its own offset; it is
//            restaged only when a block's centre differs from the staged one (median predictors of one MB mostly
//            share their integer centre)
//   task     (column x, run of 4 or 8 candidate rows): a reference row is loaded once and meets that many rows of
//            the block, which sits in registers (template on the block shape: everything unrolled)
//   argmin   packed (cost + bias) << 15 | key in 32 bits as in me_int_tb.cu (Gen A cost domain, SAD); 64-bit
//            (cost, key) for the scaled-up domain and SSE; per-block bit tables instead of a clz per candidate
// Measured at 1080p, R = 32, median predictors (tools/time_closed_loop.py): 17.8 ms for the first form (thread =
// candidate, unaligned words through L1 — kept below for search ranges shorter than a run), 3.9 ms with a window per
// block, 3.0 ms with the shared MB window = 6.0 T lane-op/s of the 530 algorithmic operations per candidate and MB
// (7 x 64 packed SADs + 41 x 2), 0.23 of the measured mix peak.
#include <cstdio>

#include "jmme_dev.cuh"

namespace {

constexpr int FK = 4;                                     // candidate rows per task (small blocks)
constexpr int FKB = 8;                                    // ... of the blocks with 16 or 8 rows

// ---- plain form: thread = candidate (R < 4) ------------------------------------------------------------------
__global__ void __launch_bounds__(128) me_full_plain_kernel(const SearchParams P)
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

// ---- window form -------------------------------------------------------------------------------------------------
struct FullLayout {
    int RS, RAWW, rows_max, off_win, off_raw, off_bits, total_words;
    __host__ __device__ explicit FullLayout(int R)
    {
        RS = 2 * R + 16;                                  // word positions 0 .. 2R + bw - 4 of the widest block
        RS |= 1;                                          // odd stride: rows of a column on different banks
        rows_max = 2 * R + 16;
        RAWW = ((15 + 2 * R + 16 + 3 + 15) & ~15) >> 2;   // raw row: the window's bytes from a 16-byte boundary on, + 3
        off_win = 0;
        off_raw = (rows_max * RS + 3) & ~3;
        off_bits = off_raw + rows_max * RAWW;
        total_words = off_bits + 4 * ((2 * R + 1 + 3) >> 2);   // bx, by of two blocks
    }
};

// all candidates of one block: tasks of (column, FK rows) over the CTA's threads; returns this thread's minimum
template <int BW4, int BH, int FK, bool WIDE>
__device__ __forceinline__ unsigned long long full_block(const SearchParams &P, const uint32_t *s_win, int RS, const uint32_t *s_cur,
                                                       const uint8_t *s_bx, const uint8_t *s_by, const uint32_t *s_T, int bx, int by,
                                                       int x00, int y00, unsigned bonus_pk, int bonus, int tid)
{
    const int ncols = P.ncols, nruns = (ncols + FK - 1) / FK;
    const bool sse = WIDE && P.metric[0] == JMME_DIST_SSE;
    uint32_t cur[BH][BW4];
#pragma unroll
    for (int r = 0; r < BH; r++)
#pragma unroll
        for (int w = 0; w < BW4; w++) cur[r][w] = s_cur[4 * (by + r) + (bx >> 2) + w];
    unsigned long long best = ~0ull;
    unsigned best32 = 0xFFFFFFFFu;
    for (int task = tid; task < ncols * nruns; task += 128) {
        const int run = task / ncols, x = task - run * ncols;
        const int y0 = min(run * FK, ncols - FK);         // the last run overlaps its predecessor (idempotent)
        unsigned acc[FK];
#pragma unroll
        for (int k = 0; k < FK; k++) acc[k] = 0;
        const uint32_t *base = s_win + y0 * RS + x;
#pragma unroll
        for (int rr = 0; rr < BH + FK - 1; rr++) {
            uint32_t rw[BW4];
#pragma unroll
            for (int w = 0; w < BW4; w++) rw[w] = base[rr * RS + 4 * w];
#pragma unroll
            for (int k = 0; k < FK; k++) {
                const int r = rr - k;
                if (r >= 0 && r < BH) {
#pragma unroll
                    for (int w = 0; w < BW4; w++) {
                        if (sse) { const unsigned d = __vabsdiffu4(cur[r][w], rw[w]); acc[k] = __dp4a(d, d, acc[k]); }
                        else acc[k] = sad4(cur[r][w], rw[w], acc[k]);
                    }
                }
            }
        }
        const unsigned bxv = s_bx[x];
#pragma unroll
        for (int k = 0; k < FK; k++) {
            const int y = y0 + k;
            const unsigned bits = bxv + s_by[y];
            const unsigned key = __ldg(P.spiral_key + y * ncols + x);
            if constexpr (!WIDE) {
                unsigned v = (acc[k] << JMME_KEY_BITS) + s_T[bits] + key;
                if (x == x00 && y == y00) v -= bonus_pk;
                best32 = min(best32, v);
            } else {
                int c = d_dscale(P.cost_domain, (int)acc[k]) + d_wcost(P.cost_domain, P.lf[0], (int)bits);
                if (x == x00 && y == y00) c -= bonus;
                const unsigned long long v = ((unsigned long long)(unsigned)(c + 0x40000000) << 32) | key;
                best = v < best ? v : best;
            }
        }
    }
    return WIDE ? best : (unsigned long long)best32;
}

template <bool WIDE>
__global__ void __launch_bounds__(128, 4) me_full_kernel(const SearchParams P)
{
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t s_cur[64];
    __shared__ unsigned long long s_best[JMME_NBLK];
    __shared__ uint32_t s_T[JMME_NT];
    const FullLayout L(P.R);
    uint32_t *s_win = smem + L.off_win;
    uint32_t *s_raw = smem + L.off_raw;
    uint8_t *s_bx = (uint8_t *)(smem + L.off_bits);
    const int bstride = (P.ncols + 3) & ~3;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = P.R, ncols = P.ncols, RS = L.RS, RAWW = L.RAWW;
    const int n_mb_stripe = d_n_units(P);
    const int n_mb = P.mb_w * P.mb_h;
    const int item = blockIdx.x;
    const int ref = item / n_mb_stripe;
    const int mb = d_unit_mb(P, item - ref * n_mb_stripe);
    const int mby = mb / P.mb_w, mbx = mb - mby * P.mb_w;
    const int npb = P.pred_policy == JMME_PRED_PER_BLOCK ? JMME_NBLK : 1;
    const int16_t *pr = P.pred ? P.pred + ((size_t)ref * n_mb + mb) * npb * 2 : nullptr;
    const int dom = P.cost_domain, lf0 = P.lf[0];
    const int bonus16 = (!P.rdopt && ref == 0) ? d_wcost(dom, lf0, 16) : 0;
    const unsigned bias = WIDE ? 0u : (unsigned)(P.rdopt ? 0 : d_weighted_cost(P.lambda_factor, 16));   // keeps cost + bias >= 0
    const uint8_t *plane = P.planes[ref];

    if (tid < 64) {
        const int row = tid >> 2, w = tid & 3;
        s_cur[tid] = *(const uint32_t *)(P.cur + (size_t)min(16 * mby + row, P.cur_h - 1) * P.cur_stride + 16 * mbx + 4 * w);
    }
    if (tid < JMME_NBLK) s_best[tid] = ~0ull;
    if (!WIDE)
        for (int i = tid; i < JMME_NT; i += 128) s_T[i] = ((unsigned)d_weighted_cost(P.lambda_factor, i) + bias) << JMME_KEY_BITS;

    // The window in shared memory is the MB's: (16 + 2R)^2 around the staged centre; a block reads it at its own offset
    // (bx, by).  It is restaged only when a block's centre differs from the staged one — median predictors of one MB
    // mostly share their integer centre, so a typical MB stages once, an MB with 41 scattered predictors 41 times.
    int scx = 0x7FFFFFFF, scy = 0x7FFFFFFF;               // staged centre (uniform over the CTA)
    const int rows = 16 + 2 * R, nwords = 16 + 2 * R - 3; // word positions the candidates of any block read
    int nblk = 0;
    for (int b = 0; b < JMME_NBLK; b++) {
        const int t = c_blk_type[b];
        if (!((P.blocktype_mask >> t) & 1)) continue;
        const int bx = c_blk_x[b], by = c_blk_y[b];
        const int px = pr ? d_pred(pr[2 * (npb == 1 ? 0 : b)]) : 0, py = pr ? d_pred(pr[2 * (npb == 1 ? 0 : b) + 1]) : 0;
        const int cx = d_clamp(px / 4, -P.cmax, P.cmax), cy = d_clamp(py / 4, -P.cmax, P.cmax);
        // this block's bit tables (double-buffered: the previous block's readers may still be at work)
        uint8_t *tbx = s_bx + (nblk & 1) * 2 * bstride, *tby = tbx + bstride;
        nblk++;
        for (int i = tid; i < ncols; i += 128) {
            tbx[i] = (uint8_t)d_se_bits(4 * (cx + i - R) - px);
            tby[i] = (uint8_t)d_se_bits(4 * (cy + i - R) - py);
        }
        if (cx != scx || cy != scy) {
            scx = cx; scy = cy;
            const int gx0 = P.pad + 16 * mbx + cx - R, gy0 = P.pad + 16 * mby + cy - R;
            const uint8_t *g = plane + (size_t)gy0 * P.pstride + (gx0 & ~15);
            const int t16 = gx0 & 15, nch = (t16 + nwords + 3 + 15) >> 4;  // 16-byte chunks per raw row
            __syncthreads();                                               // the previous window's readers are done
            for (int i = tid; i < rows * nch; i += 128) {
                const int row = i / nch, c = i - row * nch;
                *(uint4 *)(s_raw + row * RAWW + 4 * c) = __ldg((const uint4 *)(g + (size_t)row * P.pstride + 16 * c));
            }
            __syncthreads();
            // expansion: word x of a row = raw bytes t16 + x .. t16 + x + 3 (two per lane and pass; warp-uniform trips)
            const int np = (nwords + 1) >> 1;
            for (int xp0 = 0; xp0 < np; xp0 += 32) {
                const int xp = xp0 + lane;
                if (xp >= np) continue;
                const int o0 = t16 + 2 * xp, i0 = o0 >> 2, b0 = o0 & 3;
                const unsigned sel0 = 0x3210u + 0x1111u * b0;
                const unsigned sel1 = b0 == 3 ? 0x3210u : 0x3210u + 0x1111u * (b0 + 1);
                for (int row = warp; row < rows; row += 4) {
                    const uint32_t *raw = s_raw + row * RAWW + i0;
                    const uint32_t a0 = raw[0], a1 = raw[1], a2 = raw[2];
                    s_win[row * RS + 2 * xp] = __byte_perm(a0, a1, sel0);
                    s_win[row * RS + 2 * xp + 1] = b0 == 3 ? __byte_perm(a1, a2, sel1) : __byte_perm(a0, a1, sel1);
                }
            }
        }
        __syncthreads();
        const uint32_t *wblk = s_win + by * RS + bx;                   // the block's window inside the MB's
        const int x00 = R - cx, y00 = R - cy;                          // window offsets of MV (0,0) (outside when |c| > R)
        const int bonus = b == 0 ? bonus16 : 0;
        const unsigned bonus_pk = (unsigned)bonus << JMME_KEY_BITS;
        unsigned long long best;
        switch (t) {
        case 1: best = full_block<4, 16, FKB, WIDE>(P, wblk, RS, s_cur, tbx, tby, s_T, bx, by, x00, y00, bonus_pk, bonus, tid); break;
        case 2: best = full_block<4, 8, FKB, WIDE>(P, wblk, RS, s_cur, tbx, tby, s_T, bx, by, x00, y00, bonus_pk, bonus, tid); break;
        case 3: best = full_block<2, 16, FKB, WIDE>(P, wblk, RS, s_cur, tbx, tby, s_T, bx, by, x00, y00, bonus_pk, bonus, tid); break;
        case 4: best = full_block<2, 8, FKB, WIDE>(P, wblk, RS, s_cur, tbx, tby, s_T, bx, by, x00, y00, bonus_pk, bonus, tid); break;
        case 5: best = full_block<2, 4, FK, WIDE>(P, wblk, RS, s_cur, tbx, tby, s_T, bx, by, x00, y00, bonus_pk, bonus, tid); break;
        case 6: best = full_block<1, 8, FKB, WIDE>(P, wblk, RS, s_cur, tbx, tby, s_T, bx, by, x00, y00, bonus_pk, bonus, tid); break;
        default: best = full_block<1, 4, FK, WIDE>(P, wblk, RS, s_cur, tbx, tby, s_T, bx, by, x00, y00, bonus_pk, bonus, tid); break;
        }
        if constexpr (WIDE) {
            for (int o = 16; o; o >>= 1) {
                const unsigned long long u = __shfl_xor_sync(0xFFFFFFFFu, best, o);
                best = u < best ? u : best;
            }
        } else {
            best = __reduce_min_sync(0xFFFFFFFFu, (unsigned)best);
        }
        if (lane == 0) atomicMin(&s_best[b], best);
    }
    __syncthreads();
    if (tid < JMME_NBLK && ((P.blocktype_mask >> c_blk_type[tid]) & 1)) {
        const unsigned long long v = s_best[tid];
        const int px = pr ? d_pred(pr[2 * (npb == 1 ? 0 : tid)]) : 0, py = pr ? d_pred(pr[2 * (npb == 1 ? 0 : tid) + 1]) : 0;
        const int cx = d_clamp(px / 4, -P.cmax, P.cmax), cy = d_clamp(py / 4, -P.cmax, P.cmax);
        const unsigned key = WIDE ? (unsigned)v : ((unsigned)v & JMME_KEY_MASK);
        int dx, dy;
        d_spiral_xy((int)key - 1, dx, dy);
        BlkRes r;
        r.mvx = (int16_t)(4 * (cx + dx));
        r.mvy = (int16_t)(4 * (cy + dy));
        r.cost = WIDE ? (int)(unsigned)(v >> 32) - 0x40000000 : (int)((unsigned)v >> JMME_KEY_BITS) - (int)bias;
        P.res[((size_t)ref * n_mb + mb) * JMME_NBLK + tid] = r;
    }
}

}  // namespace

cudaError_t jmme_launch_me_full(const SearchParams &P, cudaStream_t st)
{
    const int n_items = d_n_units(P) * P.num_refs;
    if (n_items <= 0) return cudaSuccess;
    if (P.ncols < FKB) {
        snprintf(jmme_kernel_name_buf(), JMME_KNAME_LEN, "me_full_plain_kernel");
        me_full_plain_kernel<<<n_items, 128, 0, st>>>(P);
        return cudaGetLastError();
    }
    const bool wide = P.cost_domain || P.metric[0] == JMME_DIST_SSE;
    const FullLayout L(P.R);
    const size_t bytes = (size_t)L.total_words * 4;
    snprintf(jmme_kernel_name_buf(), JMME_KNAME_LEN, "me_full_kernel<WIDE=%d>", (int)wide);
    cudaError_t e;
    int occ = 0;
    if (wide) {
        static KernelState ks;
        e = jmme_kernel_occupancy(me_full_kernel<true>, ks, 128, bytes, &occ);
        if (e != cudaSuccess) return e;
        me_full_kernel<true><<<n_items, 128, bytes, st>>>(P);
    } else {
        static KernelState ks;
        e = jmme_kernel_occupancy(me_full_kernel<false>, ks, 128, bytes, &occ);
        if (e != cudaSuccess) return e;
        me_full_kernel<false><<<n_items, 128, bytes, st>>>(P);
    }
    return cudaGetLastError();
}
