// me_int.cu — integer full-pel search over all 41 blocks of a macroblock (sm_100a).
//
// Stands in for JM's SetupFastFullPelSearch + SetupLargerBlocks + FastFullPelBlockMotionSearch
// (and a8 when every block shares one predictor).  Nothing of this
// exists under /root/reference (README.md:1-4 only); conventions are DESIGN.md §2.
//
// Mapping
//   CTA        persistent; one (reference, macroblock) work item at a time, grid-stride order
//   pipeline   while item i is searched, the raw window rows and the current MB of item i+1 arrive
//              in shared memory through cp.async (LDGSTS, 16-byte chunks); after the search a short
//              shared-to-shared pass expands them into the search layout
//   window     (2R+16) x (2R+16) reference bytes around the search centre, held as one 32-bit word
//              PER BYTE POSITION: word[row][x] = bytes x..x+3 of the row.  A candidate at horizontal
//              offset x reads words x, x+4, x+8, x+12: always aligned, no PRMT before VABSDIFF4, and
//              32 lanes with consecutive x hit 32 consecutive banks
//   thread     a vertical run of K candidates (same dx, dy..dy+K-1): each reference row is loaded
//              once (4 LDS.32) and reused by the K candidates; the current MB (64 words) and the
//              K*16 4x4 partial SADs live in registers
//   warp       32 consecutive dx of one run (bank-conflict free); the columns left over after the
//              last full group of 32 are gathered, several runs per warp, into residual tasks
//   SAD        4x4 SADs once (64 VABSDIFF4.U8.ACC per candidate), 25 adds build the 8x4, 4x8, 8x8,
//              16x8, 8x16, 16x16 sums (SetupLargerBlocks)
//   argmin     cost and tie-break key are packed as ((sad + rate + bias) << 15) | key with
//              key = 1 + spiral index, so that ONE unsigned min per (block, candidate) reproduces
//              JM's strict-< scan in spiral order; candidates are folded in pairs (VIMNMX3);
//              41 running minima per thread, then CREDUX.MIN per warp and shared atomicMin per CTA
//   MV (0,0)   JM's "(0,0) first" pre-test (!rdopt) is the key table entry of that candidate set
//              to 0 for the item; the 16x16 (0,0) bonus is one extra 16x16 evaluation by warp 0
#include <cstdio>
#include <type_traits>

#include "jmme_dev.cuh"

namespace {

template <bool PER_BLOCK>
struct SmemLayout {
    int RS, rows, RAWW;                 // window row stride (words, multiple of 4), rows, raw row words
    int off_win, off_raw, off_cur, off_T, off_best, off_key, off_bx, off_by, total_words;
    __host__ __device__ SmemLayout(int R)
    {
        const int ncols = 2 * R + 1;
        RS = (2 * R + 13 + 3) & ~3;     // word positions 0 .. 2R+12
        rows = 2 * R + 16;
        RAWW = ((15 + RS + 12 + 15) & ~15) >> 2;      // 16-byte chunks covering any 16-byte phase
        off_win = 0;
        off_raw = off_win + rows * RS;
        off_cur = off_raw + rows * RAWW;
        off_T = off_cur + 2 * 64;
        off_best = off_T + JMME_NT;
        off_key = off_best + 96;        // 48 minima; 64-bit ones in the wide kernels
        off_bx = off_key + (ncols * ncols + 1) / 2;
        const int nb = PER_BLOCK ? JMME_NBLK : 1;
        off_by = off_bx + (nb * ncols + 3) / 4;
        total_words = off_by + (nb * ncols + 3) / 4;
    }
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

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

struct Item {
    int ref, mbx, mby, mb, cx, cy;
};

// distortion of 4 pixels accumulated: SAD (one VABSDIFF4.ACC) or SSE (VABSDIFF4, then IDP.4A of the bytes with themselves)
template <bool SSE>
__device__ __forceinline__ unsigned dist4_t(unsigned a, unsigned b, unsigned c)
{
    if constexpr (SSE) { const unsigned d = __vabsdiffu4(a, b); return __dp4a(d, d, c); }
    else return sad4(a, b, c);
}

// RS_CT: compile-time window row stride (0 = take it from the layout at run time)
// MODE 0: packed 32-bit (cost, key) minima — cost domain 0 with SAD, where cost + bias < 2^17.
// MODE 1 / 2 ("wide"): 64-bit minima (cost << 15 | key) for what does not fit 17 bits: the scaled-up cost domain
// (cost_domain = 1: (SAD << 5) + lambda_factor * bits, 22 bits) and SSE distortion (MODE 2: VABSDIFF4 then
// IDP.4A of the difference bytes with themselves; 25 bits, 30 in domain 1).  A 64-bit unsigned minimum is two
// ISETP and two SEL instead of one VIMNMX; the warp reduction splits it into two CREDUX (cost, then the key
// among the lanes that hold that cost).
template <int K, int NW, int MINB, bool PER_BLOCK, bool ONLY16, int RS_CT, int MODE = 0>
__global__ void __launch_bounds__(NW * 32, MINB) me_int_kernel(const SearchParams P)
{
    constexpr bool WIDE = MODE != 0, SSE = MODE == 2;
    using best_t = typename std::conditional<WIDE, unsigned long long, uint32_t>::type;
    extern __shared__ __align__(16) uint32_t smem[];
    const SmemLayout<PER_BLOCK> L(P.R);
    uint32_t *s_win = smem + L.off_win;
    uint32_t *s_raw = smem + L.off_raw;
    uint32_t *s_cur2 = smem + L.off_cur;
    uint32_t *s_T = smem + L.off_T;
    best_t *s_best = (best_t *)(smem + L.off_best);
    uint16_t *s_key = (uint16_t *)(smem + L.off_key);
    uint8_t *s_bx = (uint8_t *)(smem + L.off_bx);
    uint8_t *s_by = (uint8_t *)(smem + L.off_by);
    constexpr best_t BEST_MAX = ~(best_t)0;
    auto dist4 = [](unsigned a, unsigned b, unsigned c) { return dist4_t<SSE>(a, b, c); };
    const int dom = WIDE ? P.cost_domain : 0, lf0 = WIDE ? P.lf[0] : P.lambda_factor;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = P.R, ncols = P.ncols, ncand = ncols * ncols;
    const int RS = RS_CT ? RS_CT : L.RS;
    const int rows = L.rows, RAWW = L.RAWW;
    const int n_mb_stripe = d_n_units(P);
    const int n_mb = P.mb_w * P.mb_h;
    const int n_items = n_mb_stripe * P.num_refs;
    constexpr int NB = ONLY16 ? 1 : JMME_NBLK;
    constexpr int NA = ONLY16 ? 4 : 16;
    constexpr int NPB = PER_BLOCK ? JMME_NBLK : 1;

    // tables that do not depend on the work item
    for (int i = tid; i < ncand; i += NW * 32) s_key[i] = P.spiral_key[i];
    const int bonus_base = P.rdopt ? 0 : d_wcost(dom, lf0, 16);
    const unsigned bias = (unsigned)bonus_base;          // keeps (cost + bias) >= 0
    for (int i = tid; i < JMME_NT; i += NW * 32)         // rate + bias, already shifted next to the key when packed in 32 bits
        s_T[i] = ((unsigned)d_wcost(dom, lf0, i) + bias) << (WIDE ? 0 : JMME_KEY_BITS);
    const bool pretest = (!P.rdopt) && P.search_mode == JMME_SEARCH_FASTFULL;
    int patched = -1;                                    // key-table entry currently forced to 0

    // task geometry: full groups of 32 columns per run, then residual columns, several runs per warp
    const int nruns = (ncols + K - 1) / K;
    const int nseg = ncols >> 5;
    const int wr = ncols - 32 * nseg;                    // 1..31 (ncols is odd)
    const int G = 32 / wr;                               // runs per residual task
    const int n_main = nruns * nseg;
    const int n_tasks = n_main + (nruns + G - 1) / G;

    auto decode_item = [&](int item, Item &it) {
        it.ref = item / n_mb_stripe;
        it.mb = d_unit_mb(P, item - it.ref * n_mb_stripe);
        it.mby = it.mb / P.mb_w;
        it.mbx = it.mb - it.mby * P.mb_w;
        const int16_t *pr = P.pred ? P.pred + ((size_t)it.ref * n_mb + it.mb) * NPB * 2 : nullptr;
        const int p16x = pr ? d_pred(pr[0]) : 0, p16y = pr ? d_pred(pr[1]) : 0;
        it.cx = d_clamp(p16x / 4, -P.cmax, P.cmax);
        it.cy = d_clamp(p16y / 4, -P.cmax, P.cmax);
    };
    // asynchronous fetch of the raw window rows (16-byte chunks) and the current MB of an item
    auto prefetch = [&](const Item &it, int buf) {
        const uint8_t *plane = P.planes[it.ref];
        const int gx0 = P.pad + 16 * it.mbx + it.cx - R, gy0 = P.pad + 16 * it.mby + it.cy - R;
        const uint8_t *g = plane + (size_t)gy0 * P.pstride + (gx0 & ~15);
        const int nch = RAWW >> 2;
        for (int i = tid; i < rows * nch; i += NW * 32) {
            const int row = i / nch, c = i - row * nch;
            cp_async16(s_raw + row * RAWW + 4 * c, g + (size_t)row * P.pstride + 16 * c);
        }
        if (tid < 16)                                    // current MB: 16 rows of 16 bytes
            cp_async16(s_cur2 + buf * 64 + 4 * tid, P.cur + (size_t)min(16 * it.mby + tid, P.cur_h - 1) * P.cur_stride + 16 * it.mbx);
        cp_async_commit();
    };
    // raw rows -> one word per byte position; per-item tables
    auto expand = [&](const Item &it) {
        const int gx0 = P.pad + 16 * it.mbx + it.cx - R;
        const int t16 = gx0 & 15, sh = (t16 & 3) * 8, w0i = t16 >> 2, nq = RS >> 2;
        for (int i = tid; i < rows * nq; i += NW * 32) {
            const int row = i / nq, q = i - row * nq;
            const uint32_t *a = s_raw + row * RAWW + w0i + q;
            const uint32_t a0 = a[0], a1 = a[1], a2 = a[2];
            const uint32_t w0 = __funnelshift_r(a0, a1, sh), w4 = __funnelshift_r(a1, a2, sh);
            uint4 v;
            v.x = w0;
            v.y = __funnelshift_r(w0, w4, 8);
            v.z = __funnelshift_r(w0, w4, 16);
            v.w = __funnelshift_r(w0, w4, 24);
            *(uint4 *)(s_win + row * RS + 4 * q) = v;
        }
        if (tid < 48) s_best[tid] = BEST_MAX;
        const int16_t *pr = P.pred ? P.pred + ((size_t)it.ref * n_mb + it.mb) * NPB * 2 : nullptr;
        for (int i = tid; i < NPB * ncols; i += NW * 32) {
            const int b = i / ncols, o = i - b * ncols;
            const int px = pr ? d_pred(pr[2 * b]) : 0, py = pr ? d_pred(pr[2 * b + 1]) : 0;
            s_bx[i] = (uint8_t)d_se_bits(4 * (it.cx + o - R) - px);
            s_by[i] = (uint8_t)d_se_bits(4 * (it.cy + o - R) - py);
        }
        const int idx00 = (R - it.cy) * ncols + (R - it.cx);
        if (pretest && tid == 0) {                       // "(0,0) first": key 0 wins every tie
            if (patched >= 0 && patched != idx00) s_key[patched] = P.spiral_key[patched];
            s_key[idx00] = 0;
        }
        if (pretest) patched = idx00;
    };

    Item cur_it, nxt_it;
    int item = blockIdx.x, buf = 0;
    if (item >= n_items) return;
    decode_item(item, cur_it);
    prefetch(cur_it, 0);
    cp_async_wait_all();
    __syncthreads();
    expand(cur_it);
    __syncthreads();

    for (; item < n_items; item += gridDim.x) {
        const int nxt = item + gridDim.x;
        const bool has_next = nxt < n_items;
        if (has_next) {
            decode_item(nxt, nxt_it);
            prefetch(nxt_it, buf ^ 1);                   // lands while this item is searched
        }
        const int cx = cur_it.cx, cy = cur_it.cy;
        const int bonus = (cur_it.ref == 0) ? bonus_base : 0;
        const int x00 = R - cx, y00 = R - cy;            // window offsets of MV (0,0)
        const uint32_t *s_cur = s_cur2 + buf * 64;

        // The current MB is the same for every lane; an opaque zero lane offset keeps ptxas from
        // parking it in uniform registers (it then pays a UR->R move per VABSDIFF4 operand).
        unsigned lz;
        asm volatile("and.b32 %0, %1, 0;" : "=r"(lz) : "r"(lane));
        uint32_t cur[16][4];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const uint4 v = *(const uint4 *)(s_cur + lz + 4 * r);
            cur[r][0] = v.x; cur[r][1] = v.y; cur[r][2] = v.z; cur[r][3] = v.w;
        }

        if (bonus != 0 && warp == 0) {
            // 16x16 block at MV (0,0) with the -WEIGHTED_COST(lambda,16) bonus: one more candidate
            // evaluation for block 0, spread over the 32 lanes (2 words each)
            unsigned s = 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int i = lane + 32 * h, row = i >> 2, j = i & 3;
                s = dist4(s_cur[i], s_win[(y00 + row) * RS + x00 + 4 * j], s);
            }
            s = __reduce_add_sync(0xFFFFFFFFu, s);
            if (lane == 0) {
                if constexpr (WIDE) {
                    const unsigned c32 = (unsigned)d_dscale(dom, (int)s) + s_T[s_bx[x00] + s_by[y00]] - (unsigned)bonus;
                    atomicMin(&s_best[0], ((best_t)c32 << JMME_KEY_BITS) | s_key[y00 * ncols + x00]);
                } else {
                    const unsigned v = (s << JMME_KEY_BITS) + s_T[s_bx[x00] + s_by[y00]] + s_key[y00 * ncols + x00] -
                                       ((unsigned)bonus << JMME_KEY_BITS);
                    atomicMin(&s_best[0], (best_t)v);
                }
            }
        }

        best_t best[NB];
#pragma unroll
        for (int b = 0; b < NB; b++) best[b] = BEST_MAX;

        for (int task = warp; task < n_tasks; task += NW) {
            int run, xoff;
            if (task < n_main) {
                run = task / nseg;
                xoff = 32 * (task - run * nseg) + lane;
            } else {
                int g = lane / wr, x = lane - g * wr;
                if (g >= G) { g = 0; x = 0; }            // idle lanes repeat lane 0 (idempotent)
                run = min((task - n_main) * G + g, nruns - 1);
                xoff = 32 * nseg + x;
            }
            const int ybase = min(run * K, ncols - K);   // last run overlaps the previous one
            const uint32_t *base = s_win + ybase * RS + xoff;

            unsigned acc[K][NA];
#pragma unroll
            for (int k = 0; k < K; k++)
#pragma unroll
                for (int i = 0; i < NA; i++) acc[k][i] = 0;
#pragma unroll
            for (int rr = 0; rr < 16 + K - 1; rr++) {
                const uint32_t *rp = base + rr * RS;
                const unsigned r0 = rp[0], r1 = rp[4], r2 = rp[8], r3 = rp[12];
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int cr = rr - k;               // current-MB row this reference row meets
                    if (cr >= 0 && cr < 16) {
                        const int a = ONLY16 ? 0 : (cr >> 2) * 4;
                        acc[k][a + 0] = dist4(cur[cr][0], r0, acc[k][a + 0]);
                        acc[k][a + 1] = dist4(cur[cr][1], r1, acc[k][a + 1]);
                        acc[k][a + 2] = dist4(cur[cr][2], r2, acc[k][a + 2]);
                        acc[k][a + 3] = dist4(cur[cr][3], r3, acc[k][a + 3]);
                    }
                }
            }

            // rate + key of candidate k (uniform predictor) or the key alone (per-block predictors); wide: the
            // rate (+ bias) in the high word, the key in the low one
            const unsigned bx0 = s_bx[xoff];
            auto kr_of = [&](int k) -> best_t {
                const int yoff = ybase + k;
                const unsigned key = s_key[yoff * ncols + xoff];
                if constexpr (WIDE) return PER_BLOCK ? (best_t)key : (((best_t)s_T[bx0 + s_by[yoff]] << JMME_KEY_BITS) | key);
                else return PER_BLOCK ? key : s_T[bx0 + s_by[yoff]] + key;
            };
            auto pack = [&](int k, best_t kr, best_t (&pk)[NB]) {
                auto one = [&](unsigned dist) -> best_t {      // distortion -> its place next to rate and key
                    if constexpr (WIDE) return (best_t)(unsigned)d_dscale(dom, (int)dist) << JMME_KEY_BITS;
                    else return dist << JMME_KEY_BITS;
                };
                if constexpr (ONLY16) {
                    const unsigned s = (acc[k][0] + acc[k][1]) + (acc[k][2] + acc[k][3]);
                    pk[0] = one(s) + kr;
                } else {
                    unsigned o[JMME_NBLK];
                    larger_blocks(acc[k], o);
                    if constexpr (!PER_BLOCK) {
#pragma unroll
                        for (int b = 0; b < NB; b++) pk[b] = one(o[b]) + kr;
                    } else {
                        const int yoff = ybase + k;
#pragma unroll
                        for (int b = 0; b < NB; b++) {
                            const unsigned rt = s_T[s_bx[b * ncols + xoff] + s_by[b * ncols + yoff]];
                            pk[b] = one(o[b]) + (WIDE ? (best_t)rt << JMME_KEY_BITS : (best_t)rt) + kr;
                        }
                    }
                }
            };
            // candidates are folded two at a time: min(best, min(a, b)) is one VIMNMX3
#pragma unroll
            for (int k = 0; k + 1 < K; k += 2) {
                best_t pa[NB], pb[NB];
                pack(k, kr_of(k), pa);
                pack(k + 1, kr_of(k + 1), pb);
#pragma unroll
                for (int b = 0; b < NB; b++) best[b] = min(best[b], min(pa[b], pb[b]));
            }
            if (K & 1) {
                best_t pa[NB];
                pack(K - 1, kr_of(K - 1), pa);
#pragma unroll
                for (int b = 0; b < NB; b++) best[b] = min(best[b], pa[b]);
            }
        }

        // ---- reduce: lanes -> warp (CREDUX.MIN), lane b keeps block b, two shared atomicMin ------
        {
            best_t m0 = BEST_MAX, m1 = BEST_MAX;
#pragma unroll
            for (int b = 0; b < NB; b++) {
                best_t m;
                if constexpr (WIDE) {            // cost first, then the smallest key among the lanes that hold that cost
                    const unsigned c = (unsigned)(best[b] >> JMME_KEY_BITS), kk = (unsigned)best[b] & JMME_KEY_MASK;
                    const unsigned mc = __reduce_min_sync(0xFFFFFFFFu, c);
                    const unsigned mk = __reduce_min_sync(0xFFFFFFFFu, c == mc ? kk : 0xFFFFFFFFu);
                    m = ((best_t)mc << JMME_KEY_BITS) | mk;
                } else {
                    m = __reduce_min_sync(0xFFFFFFFFu, best[b]);
                }
                if (b < 32) m0 = (lane == b) ? m : m0;
                else m1 = (lane == b - 32) ? m : m1;
            }
            if (lane < NB) atomicMin(&s_best[lane], m0);
            if (NB > 32 && lane < NB - 32) atomicMin(&s_best[32 + lane], m1);
        }
        cp_async_wait_all();                             // next item's raw rows have landed
        __syncthreads();                                 // every warp is done with s_win / s_best
        if (tid < NB) {
            const best_t v = s_best[tid];
            const unsigned key = (unsigned)v & JMME_KEY_MASK;
            int mvx = 0, mvy = 0;                        // key 0 = the MV (0,0) pre-test
            if (key) {
                mvx = cx + P.spiral_xy[2 * (key - 1)];
                mvy = cy + P.spiral_xy[2 * (key - 1) + 1];
            }
            BlkRes r;
            r.mvx = (int16_t)(4 * mvx);
            r.mvy = (int16_t)(4 * mvy);
            r.cost = (int)(v >> JMME_KEY_BITS) - (int)bias;
            P.res[((size_t)cur_it.ref * n_mb + cur_it.mb) * JMME_NBLK + tid] = r;
        }
        if (!has_next) break;
        __syncthreads();                                 // s_best read before expand() resets it
        expand(nxt_it);
        __syncthreads();
        cur_it = nxt_it;
        buf ^= 1;
    }
}

template <int K, int NW, int MINB, bool PER_BLOCK, bool ONLY16, int RS_CT, int MODE = 0>
cudaError_t launch_one(const SearchParams &P, int num_sms, cudaStream_t st)
{
    SmemLayout<PER_BLOCK> L(P.R);
    size_t bytes = (size_t)L.total_words * 4;
    auto kern = me_int_kernel<K, NW, MINB, PER_BLOCK, ONLY16, RS_CT, MODE>;
    static KernelState ks;                               // shared-memory opt-in and occupancy, per device
    int c_occ = 0;
    cudaError_t e = jmme_kernel_occupancy(kern, ks, NW * 32, bytes, &c_occ);
    if (e != cudaSuccess) return e;
    snprintf(jmme_kernel_name_buf(), JMME_KNAME_LEN, "me_int_kernel<K=%d,NW=%d,MINB=%d,PER_BLOCK=%d,ONLY16=%d,RS_CT=%d,MODE=%d>", K, NW,
             MINB, (int)PER_BLOCK, (int)ONLY16, RS_CT, MODE);
    int n_items = d_n_units(P) * P.num_refs;
    int grid = min(n_items, num_sms * c_occ);
    kern<<<grid, NW * 32, bytes, st>>>(P);
    return cudaGetLastError();
}

template <int K, int NW, int MINB>
cudaError_t launch_shape(const SearchParams &P, int num_sms, cudaStream_t st)
{
    const bool per_block = P.pred_policy == JMME_PRED_PER_BLOCK;
    const bool only16 = P.blocktype_mask == JMME_MASK_16x16;
    if (only16 && !per_block) return launch_one<K, NW, MINB, false, true, 0>(P, num_sms, st);
    if (per_block) return launch_one<K, NW, MINB, true, false, 0>(P, num_sms, st);
    if (P.R == 32) return launch_one<K, NW, MINB, false, false, 80>(P, num_sms, st);   // RS of R=32
    if (P.R == 64) return launch_one<K, NW, MINB, false, false, 144>(P, num_sms, st);  // RS of R=64
    return launch_one<K, NW, MINB, false, false, 0>(P, num_sms, st);
}

}  // namespace

// variant = 10*K + c:  K candidates per thread run (must be <= 2R+1);
//   c = 0: 6 warps, >= 2 CTAs/SM (<= 168 registers)    c = 1: 8 warps, 1 CTA/SM (<= 255 registers)
//   c = 2: 4 warps, >= 3 CTAs/SM (<= 168 registers)
cudaError_t jmme_launch_me_int_tb(const SearchParams &P, int num_sms, int K, int shape, cudaStream_t st);

// the wide kernels (cost domain 1, SSE): one shape, 8 warps and one CTA per SM (the 64-bit minima double the registers)
template <int MODE>
cudaError_t launch_wide(const SearchParams &P, int num_sms, cudaStream_t st)
{
    const bool per_block = P.pred_policy == JMME_PRED_PER_BLOCK;
    if (P.blocktype_mask == JMME_MASK_16x16 && !per_block) return launch_one<2, 8, 1, false, true, 0, MODE>(P, num_sms, st);
    if (P.ncols < 2) return cudaErrorInvalidValue;
    if (per_block) return launch_one<2, 8, 1, true, false, 0, MODE>(P, num_sms, st);
    return launch_one<2, 8, 1, false, false, 0, MODE>(P, num_sms, st);
}

cudaError_t jmme_launch_me_int(const SearchParams &P, int num_sms, int variant, cudaStream_t st)
{
    if (P.metric[0] == JMME_DIST_SSE) return launch_wide<2>(P, num_sms, st);
    if (P.cost_domain) return launch_wide<1>(P, num_sms, st);
    // default: measured best per search range; a wavefront step has fewer MBs than SMs, so
    // it takes the widest CTA (12 warps per MB)
    if (variant <= 0) variant = P.R <= 32 ? (P.mb_list ? 64 : 68) : 51;
    int K = variant / 10, c = variant % 10;
    if (c >= 4) {                                   // two-threads-per-candidate kernel (me_int_tb.cu)
        if (P.blocktype_mask != JMME_MASK_16x16 && K <= P.ncols) return jmme_launch_me_int_tb(P, num_sms, K, c, st);
        K = 3; c = 2;                               // 16x16 only, or a window narrower than one run
    }
    if (K > P.ncols) K = 3;
#define PICKC(KK, CC, NWW, MB) \
    if (K == KK && c == CC) return launch_shape<KK, NWW, MB>(P, num_sms, st);
    PICKC(3, 0, 6, 2)
    PICKC(3, 1, 8, 1) PICKC(5, 1, 8, 1)
    PICKC(2, 2, 4, 3) PICKC(3, 2, 4, 3)
#undef PICKC
    return cudaErrorInvalidValue;
}
