// me_int.cu — integer full-pel search over all 41 blocks of a macroblock (sm_100a).
//
// Stands in for JM's SetupFastFullPelSearch + SetupLargerBlocks + FastFullPelBlockMotionSearch
// (and a8 when every block shares one predictor).  Nothing of this
// exists under /root/reference (README.md:1-4 only); conventions are DESIGN.md §2.
//
// Mapping
//   CTA        one (reference, macroblock) work item at a time, persistent grid-stride loop
//   window     (2R+16) x (2R+16) reference bytes around the search centre, staged in shared
//              memory as FOUR byte-phase copies (copy p = window shifted left by p bytes), so
//              that every candidate reads aligned 32-bit words and VABSDIFF4 needs no PRMT
//   thread     a vertical run of K candidates (same dx, dy..dy+K-1): each reference row is
//              loaded once (4 LDS.32) and reused by the K candidates; the current MB (64 words)
//              and the K*16 4x4 partial SADs live in registers
//   SAD        4x4 SADs once (64 VABSDIFF4.U8.ACC per candidate), 25 adds build the 8x4, 4x8,
//              8x8, 16x8, 8x16, 16x16 sums (SetupLargerBlocks)
//   argmin     cost and tie-break key are packed as ((sad + rate + bias) << 15) | key with
//              key = 1 + spiral index (0 for the MV(0,0) pre-test), so that ONE unsigned min per
//              (block, candidate) reproduces JM's strict-< scan in spiral order; 41 running
//              minima per thread, then CREDUX.MIN per warp and shared-memory atomicMin per CTA
#include "jmme_dev.cuh"

namespace {

template <bool PER_BLOCK>
struct SmemLayout {
    int RS, rows, copy_stride;          // words
    int off_win, off_cur, off_T, off_best, off_key, off_bx, off_by, total_words;
    __host__ __device__ SmemLayout(int R)
    {
        int ncols = 2 * R + 1;
        RS = ((2 * R) >> 2) + 4;
        rows = 2 * R + 16;
        copy_stride = rows * RS;
        copy_stride += (8 - (copy_stride & 31) + 32) & 31;     // == 8 (mod 32): conflict-free phases
        off_win = 0;
        off_cur = off_win + 4 * copy_stride;
        off_T = off_cur + 64;
        off_best = off_T + JMME_NT;
        off_key = off_best + 48;
        off_bx = off_key + (ncols * ncols + 1) / 2;
        int nb = PER_BLOCK ? JMME_NBLK : 1;
        off_by = off_bx + (nb * ncols + 3) / 4;
        total_words = off_by + (nb * ncols + 3) / 4;
    }
};

// partition sums of one candidate from its 16 4x4 SADs (SetupLargerBlocks), result order
__device__ __forceinline__ void larger_blocks(const unsigned (&s)[16], unsigned (&o)[JMME_NBLK])
{
#pragma unroll
    for (int i = 0; i < 16; i++) o[25 + i] = s[i];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) o[17 + 4 * j + i] = s[8 * j + i] + s[8 * j + 4 + i];       // 4x8
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 2; i++) o[9 + 2 * j + i] = s[4 * j + 2 * i] + s[4 * j + 2 * i + 1]; // 8x4
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int i = 0; i < 2; i++) o[5 + 2 * j + i] = o[9 + 4 * j + i] + o[9 + 4 * j + 2 + i]; // 8x8
    o[3] = o[5] + o[7];                                                                          // 8x16
    o[4] = o[6] + o[8];
    o[1] = o[5] + o[6];                                                                          // 16x8
    o[2] = o[7] + o[8];
    o[0] = o[1] + o[2];                                                                          // 16x16
}

template <int K, int NW, int MINB, bool PER_BLOCK, bool ONLY16>
__global__ void __launch_bounds__(NW * 32, MINB) me_int_kernel(const SearchParams P)
{
    extern __shared__ uint32_t smem[];
    const SmemLayout<PER_BLOCK> L(P.R);
    uint32_t *s_win = smem + L.off_win;
    uint32_t *s_cur = smem + L.off_cur;
    uint32_t *s_T = smem + L.off_T;
    uint32_t *s_best = smem + L.off_best;
    uint16_t *s_key = (uint16_t *)(smem + L.off_key);
    uint8_t *s_bx = (uint8_t *)(smem + L.off_bx);
    uint8_t *s_by = (uint8_t *)(smem + L.off_by);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = P.R, ncols = P.ncols, ncand = ncols * ncols;
    const int RS = L.RS, rows = L.rows, cstride = L.copy_stride;
    const int n_mb_stripe = (P.mb_row_end - P.mb_row_begin) * P.mb_w;
    const int n_mb = P.mb_w * P.mb_h;
    const int n_items = n_mb_stripe * P.num_refs;
    constexpr int NB = ONLY16 ? 1 : JMME_NBLK;

    // tables that do not depend on the work item
    for (int i = tid; i < ncand; i += NW * 32) s_key[i] = P.spiral_key[i];
    const int bonus_base = P.rdopt ? 0 : d_weighted_cost(P.lambda_factor, 16);
    const unsigned bias = (unsigned)bonus_base;          // keeps (cost + bias) >= 0
    for (int i = tid; i < JMME_NT; i += NW * 32)
        s_T[i] = ((unsigned)d_weighted_cost(P.lambda_factor, i) + bias) << JMME_KEY_BITS;

    const int nruns = (ncols + K - 1) / K;
    const int W = nruns * ncols;

    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ref = item / n_mb_stripe;
        const int mbi = item - ref * n_mb_stripe;
        const int mby = P.mb_row_begin + mbi / P.mb_w;
        const int mbx = mbi % P.mb_w;
        const int mb = mby * P.mb_w + mbx;
        const int npb = PER_BLOCK ? JMME_NBLK : 1;
        const int16_t *pr = P.pred ? P.pred + ((size_t)ref * n_mb + mb) * npb * 2 : nullptr;
        const int p16x = pr ? pr[0] : 0, p16y = pr ? pr[1] : 0;
        const int cx = d_clamp(p16x / 4, -R, R), cy = d_clamp(p16y / 4, -R, R);
        const int bonus = (ref == 0) ? bonus_base : 0;
        const bool pretest = (!P.rdopt) && P.search_mode == JMME_SEARCH_FASTFULL;
        const bool special = pretest || bonus != 0;
        const int x00 = R - cx, y00 = R - cy;            // window offsets of MV (0,0)

        __syncthreads();                                 // previous item finished with smem
        // ---- stage the search window as 4 byte-phase copies --------------------------------
        {
            const uint8_t *plane = P.planes[ref];
            const int gx0 = P.pad + 16 * mbx + cx - R;   // byte column of window column 0
            const int gy0 = P.pad + 16 * mby + cy - R;
            const int t = gx0 & 3;
            const uint32_t *g32 = (const uint32_t *)(plane + (size_t)gy0 * P.pstride + (gx0 & ~3));
            const int pw = P.pstride >> 2;
            for (int idx = tid; idx < rows * RS; idx += NW * 32) {
                int row = idx / RS, w = idx - row * RS;
                const uint32_t *g = g32 + (size_t)row * pw + w;
                uint32_t a0 = __ldg(g), a1 = __ldg(g + 1), a2 = __ldg(g + 2);
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    int o = (t + p) >> 2, sh = ((t + p) & 3) * 8;
                    uint32_t lo = o ? a1 : a0, hi = o ? a2 : a1;
                    s_win[p * cstride + idx] = __funnelshift_r(lo, hi, sh);
                }
            }
            // current macroblock: 16 rows x 4 words
            if (tid < 64) {
                int row = tid >> 2, w = tid & 3;
                s_cur[tid] = *(const uint32_t *)(P.cur + (size_t)(16 * mby + row) * P.cur_stride + 16 * mbx + 4 * w);
            }
            if (tid < 48) s_best[tid] = 0xFFFFFFFFu;
            // MV-bit tables of this item: bits of (4*mv - pred) per window column / row
            const int nb = PER_BLOCK ? JMME_NBLK : 1;
            for (int i = tid; i < nb * ncols; i += NW * 32) {
                int b = i / ncols, o = i - b * ncols;
                int px = pr ? pr[2 * b] : 0, py = pr ? pr[2 * b + 1] : 0;
                s_bx[i] = (uint8_t)d_se_bits(4 * (cx + o - R) - px);
                s_by[i] = (uint8_t)d_se_bits(4 * (cy + o - R) - py);
            }
        }
        __syncthreads();

        uint32_t cur[16][4];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            uint4 v = *(const uint4 *)(s_cur + 4 * r);
            cur[r][0] = v.x; cur[r][1] = v.y; cur[r][2] = v.z; cur[r][3] = v.w;
        }
        uint32_t best[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) best[b] = 0xFFFFFFFFu;

        for (int it0 = warp * 32; it0 < W; it0 += NW * 32) {
            const int idx = min(it0 + lane, W - 1);      // idle lanes repeat the last item (idempotent)
            const int run = idx / ncols;
            const int xoff = idx - run * ncols;
            const int ybase = min(run * K, ncols - K);   // last run overlaps the previous one
            const uint32_t *base = s_win + (xoff & 3) * cstride + ybase * RS + (xoff >> 2);

            constexpr int NA = ONLY16 ? 4 : 16;
            unsigned acc[K][NA];
#pragma unroll
            for (int k = 0; k < K; k++)
#pragma unroll
                for (int i = 0; i < NA; i++) acc[k][i] = 0;
#pragma unroll
            for (int rr = 0; rr < 16 + K - 1; rr++) {
                const uint32_t *rp = base + rr * RS;
                const unsigned r0 = rp[0], r1 = rp[1], r2 = rp[2], r3 = rp[3];
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int cr = rr - k;               // current-MB row this reference row meets
                    if (cr >= 0 && cr < 16) {
                        const int a = ONLY16 ? 0 : (cr >> 2) * 4;
                        acc[k][a + 0] = sad4(cur[cr][0], r0, acc[k][a + 0]);
                        acc[k][a + 1] = sad4(cur[cr][1], r1, acc[k][a + 1]);
                        acc[k][a + 2] = sad4(cur[cr][2], r2, acc[k][a + 2]);
                        acc[k][a + 3] = sad4(cur[cr][3], r3, acc[k][a + 3]);
                    }
                }
            }

            const unsigned bx0 = s_bx[xoff];
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int yoff = ybase + k;
                const unsigned key = s_key[yoff * ncols + xoff];
                unsigned pk[NB];
                if constexpr (ONLY16) {
                    unsigned s = (acc[k][0] + acc[k][1]) + (acc[k][2] + acc[k][3]);
                    pk[0] = (s << JMME_KEY_BITS) + s_T[bx0 + s_by[yoff]] + key;
                } else {
                    unsigned o[JMME_NBLK];
                    larger_blocks(acc[k], o);
                    if constexpr (!PER_BLOCK) {
                        const unsigned kr = s_T[bx0 + s_by[yoff]] + key;
#pragma unroll
                        for (int b = 0; b < NB; b++) pk[b] = (o[b] << JMME_KEY_BITS) + kr;
                    } else {
#pragma unroll
                        for (int b = 0; b < NB; b++) {
                            const unsigned kr = s_T[s_bx[b * ncols + xoff] + s_by[b * ncols + yoff]] + key;
                            pk[b] = (o[b] << JMME_KEY_BITS) + kr;
                        }
                    }
                }
#pragma unroll
                for (int b = 0; b < NB; b++) best[b] = min(best[b], pk[b]);
                if (special && xoff == x00 && yoff == y00) {
                    // the MV (0,0) candidate: tested first when !rdopt (key 0 wins every tie) and the
                    // 16x16 block gets the -WEIGHTED_COST(lambda,16) bonus on reference 0
#pragma unroll
                    for (int b = 0; b < NB; b++) {
                        unsigned v = pk[b];
                        if (pretest) v &= ~JMME_KEY_MASK;
                        if (b == 0) v -= (unsigned)bonus << JMME_KEY_BITS;
                        best[b] = min(best[b], v);
                    }
                }
            }
        }

        // ---- reduce: lanes -> warp (CREDUX.MIN) -> CTA (shared atomicMin) ---------------------
#pragma unroll
        for (int b = 0; b < NB; b++) {
            unsigned m = __reduce_min_sync(0xFFFFFFFFu, best[b]);
            if (lane == 0) atomicMin(&s_best[b], m);
        }
        __syncthreads();
        if (tid < NB) {
            const unsigned v = s_best[tid];
            const unsigned key = v & JMME_KEY_MASK;
            int mvx = 0, mvy = 0;
            if (key) {
                mvx = cx + P.spiral_xy[2 * (key - 1)];
                mvy = cy + P.spiral_xy[2 * (key - 1) + 1];
            }
            BlkRes r;
            r.mvx = (int16_t)(4 * mvx);
            r.mvy = (int16_t)(4 * mvy);
            r.cost = (int)(v >> JMME_KEY_BITS) - (int)bias;
            P.res[((size_t)ref * n_mb + mb) * JMME_NBLK + tid] = r;
        }
    }
}

template <int K, int NW, int MINB, bool PER_BLOCK, bool ONLY16>
cudaError_t launch_one(const SearchParams &P, int num_sms, cudaStream_t st)
{
    SmemLayout<PER_BLOCK> L(P.R);
    size_t bytes = (size_t)L.total_words * 4;
    auto kern = me_int_kernel<K, NW, MINB, PER_BLOCK, ONLY16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, bytes);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    int n_items = (P.mb_row_end - P.mb_row_begin) * P.mb_w * P.num_refs;
    int grid = min(n_items, num_sms * occ);
    kern<<<grid, NW * 32, bytes, st>>>(P);
    return cudaGetLastError();
}

}  // namespace

// K: candidates per thread run; chosen by the host (tuning knob), must be <= 2R+1
cudaError_t jmme_launch_me_int(const SearchParams &P, int num_sms, int K, cudaStream_t st)
{
    const bool per_block = P.pred_policy == JMME_PRED_PER_BLOCK;
    const bool only16 = P.blocktype_mask == JMME_MASK_16x16;
    if (K > P.ncols) K = 3;
// K <= 3: 6 warps, 2 CTAs/SM (<=168 registers); K >= 4: 8 warps, 1 CTA/SM (<=255 registers)
#define GO(KK, PB, O16) return launch_one<KK, (KK <= 3 ? 6 : 8), (KK <= 3 ? 2 : 1), PB, O16>(P, num_sms, st)
#define PICK(KK)                                  \
    if (K == KK) {                                \
        if (only16 && !per_block) GO(KK, false, true); \
        if (per_block) GO(KK, true, false);       \
        GO(KK, false, false);                     \
    }
    PICK(2) PICK(3) PICK(4) PICK(5)
#undef PICK
#undef GO
    return cudaErrorInvalidValue;
}
