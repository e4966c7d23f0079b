// me_int_tb.cu — integer full-pel search, all 41 blocks, "two threads per candidate" mapping (sm_100a).
//
// Same function as me_int.cu (JM SetupFastFullPelSearch + SetupLargerBlocks +
// FastFullPelBlockMotionSearch) and the same pipeline, window layout and
// packed (cost,key) argmin; what changes is the thread mapping, chosen from the ncu profile of
// me_int.cu (profiles/r01_*): one thread per candidate needs 64 (current MB) + 16K (4x4 SADs) +
// 41 (minima) live registers, which at 12 warps/SM (168 registers) made ptxas park the current MB
// in uniform registers and pay one UR->R move per ~1.4 VABSDIFF4.
//
//   thread pair  lanes l and l+16 share a run of K candidates: lane l (T) owns MB rows 0-7, lane
//                l+16 (B) rows 8-15.  Each holds 32 words of the current MB, 8K partial SADs and 22
//                minima (19 blocks inside its half + 16x16, 8x16 left, 8x16 right, for which the
//                halves exchange their two 8x8 sums with SHFL.BFLY)
//   banks        the window row stride is == 2 (mod 4) words, so the B lanes (8 rows further down)
//                sit 16 banks away from the T lanes: 16 + 16 consecutive banks, conflict-free
//   K            up to 8 candidates per run: 4*(7+K)/K row words per thread and candidate
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>

#include <cooperative_groups.h>

#include "jmme_dev.cuh"
#include "wave.cuh"

namespace {

struct TbLayout {
    int RS, rows, RAWW;                 // window row stride (words, == 2 mod 4), rows, raw row words
    int off_win, off_raw, off_cur, off_T, off_best, off_task, off_key, off_bx, off_by, off_kr, total_words;
    // nmb: the item is nmb horizontally adjacent MBs sharing one (16*(nmb-1) columns wider) window
    __host__ __device__ TbLayout(int R, bool per_block, bool key_in_smem, int K, bool kr_table = false, int nmb = 1)
    {
        const int ncols = 2 * R + 1;
        RS = 2 * R + 13 + 16 * (nmb - 1);           // word positions 0 .. 2R+12 (+16 per further MB)
        while ((RS & 3) != 2) RS++;
        rows = 2 * R + 16;
        RAWW = ((15 + RS + 12 + 15) & ~15) >> 2;
        off_win = 0;
        off_raw = (off_win + rows * RS + 3) & ~3;
        off_cur = off_raw + rows * RAWW;
        off_T = off_cur + 2 * 64 * nmb;
        off_best = off_T + JMME_NT;
        off_task = off_best + 48 * nmb;                       // u16 (ybase << 8 | xbase) per main task
        off_key = (off_task + ((((ncols + K - 1) / K) * (ncols >> 4) + 1) >> 1) + 3) & ~3;   // 16-byte aligned (uint4 copy)
        off_bx = (off_key + (key_in_smem ? (ncols * ncols + 1) / 2 : 0) + 1) & ~1;      // 8-byte aligned (LDS.64)
        // one predictor: bits per column / row (bytes).  41 predictors: per (half, column|row) the 22 local
        // blocks' bits * 4, padded to 24 bytes, so that a thread fetches all of them as six words
        const int tw = per_block ? (2 * ncols * 24) / 4 : (ncols + 3) / 4;
        off_by = (off_bx + tw + 1) & ~1;
        off_kr = (off_by + tw + 3) & ~3;   // rate+key of every candidate (zero predictors only), 16-byte aligned
        total_words = off_kr + (kr_table ? ncols * ncols : 0);
    }
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// local block order of one half: [0] 16x8, [1,2] 8x8, [3..6] 8x4, [7..10] 4x8, [11..18] 4x4,
// [19] 16x16, [20] 8x16 left, [21] 8x16 right;  global block = kGT[l] + half * kDL[l]
constexpr int NL = 22;
__device__ constexpr int kGT[NL] = {1, 5, 6, 9, 10, 11, 12, 17, 18, 19, 20, 25, 26, 27, 28, 29, 30, 31, 32, 0, 3, 4};
__device__ constexpr int kDL[NL] = {1, 2, 2, 4, 4, 4, 4, 4, 4, 4, 4, 8, 8, 8, 8, 8, 8, 8, 8, 0, 0, 0};

struct Item {
    int ref, mbx, mby, mb, cx, cy, nmb;
};

// RS_CT: compile-time window row stride (0 = from the layout at run time): row addresses become immediates
// KEYG: read the spiral keys from global memory (L1) instead of a shared-memory copy — frees 33 KB at R = 64
// so that two CTAs fit on an SM
// KRTAB: zero predictors (P.pred == nullptr): centre and rate are the same for every item, so rate + key of
// all (2R+1)^2 candidates is tabulated once per CTA and a candidate costs one LDS instead of three
// NMB > 1 (with KRTAB): an item is NMB horizontally adjacent MBs.  Their windows overlap in all but 16 columns
// each, so one prefetch + expansion (1.6x the work of one for NMB = 4) and one set of barriers serves all of
// them, and the NMB x 45 tasks split evenly over the warps.
// CL > 1 (MB lists of the in-frame median wavefront, where a step has fewer MBs than the GPU has SMs): a
// thread-block cluster of CL CTAs works on one item; every CTA stages the window, takes every CL-th group of
// tasks, and the partial minima meet in CTA 0's shared memory (distributed shared memory atomicMin).
// WP (with PER_BLOCK and an MB list): the kernel computes the 41 median predictors of its MB from the committed
// field itself (wave.cuh) instead of reading P.pred, which it fills for the sub-pel kernel.
// LIN (with PER_BLOCK): lambda is an integer (every !rdopt lambda is), so the rate is linear in the bits and
// one IMAD replaces the table look-up: rate << 15 = (4 bits) * (lambda * 8192).
// BAL (with KRTAB, zero predictors): the stripe's tasks — (2R+1)^2 candidates of an MB = n_tasks runs of K x 16 — are dealt
// out evenly: CTA c takes the task range [c G / n, (c + 1) G / n) of the G = items x NMB x n_tasks tasks, i.e. a partial
// first item, whole items, a partial last item.  Every CTA works the same number of tasks (+-1) whatever the stripe
// size, so there is no partial last round (1080p: 9.19 rounds of 444 CTAs used to cost 10; a 9-row stripe of an 8-GPU
// run 1.22 rounds cost 2).  An MB whose tasks are split over CTAs gets its minima combined by atomicMin on the packed
// (cost, key) words in global memory (P.gbest) — the same operation as inside a CTA, so the result does not depend
// on the partition; the kernel that consumes the integer result (sub-pel or reference selection) decodes the words
// and resets them to 0xFFFFFFFF for the next search (d_take_packed).
template <int K, int NW, int MINB, bool PER_BLOCK, int RS_CT, bool KEYG, bool KRTAB, int NMB, int CL, bool WP, bool LIN, bool BAL>
__global__ void __launch_bounds__(NW * 32, MINB) me_int_tb_kernel(const SearchParams P)
{
    __shared__ WaveNb s_wnb[WP ? 10 : 1];
    __shared__ int16_t s_wpred[WP ? 2 * JMME_NBLK : 2];
    __shared__ uint8_t s_wsrc[WP ? JMME_NBLK : 1][4], s_wtp[WP ? JMME_NBLK : 1];
    extern __shared__ __align__(16) uint32_t smem[];
    const unsigned crank = CL > 1 ? cooperative_groups::this_cluster().block_rank() : 0u;
    const TbLayout L(P.R, PER_BLOCK, !KEYG && !KRTAB, K, KRTAB, NMB);
    constexpr int NM = NMB, CURW = 64 * NM;     // MBs per item, words of one current-MB stage
    uint32_t *s_win = smem + L.off_win;
    uint32_t *s_raw = smem + L.off_raw;
    uint32_t *s_cur2 = smem + L.off_cur;
    uint32_t *s_T = smem + L.off_T;
    uint32_t *s_best = smem + L.off_best;
    uint16_t *s_task = (uint16_t *)(smem + L.off_task);
    uint16_t *s_key = (uint16_t *)(smem + L.off_key);
    uint8_t *s_bx = (uint8_t *)(smem + L.off_bx);
    uint8_t *s_by = (uint8_t *)(smem + L.off_by);
    uint32_t *s_kr = smem + L.off_kr;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, l16 = lane & 15;
    const int R = P.R, ncols = P.ncols, ncand = ncols * ncols;
    const int RS = RS_CT ? RS_CT : L.RS;
    const int rows = L.rows, RAWW = L.RAWW;
    const int n_mb = P.mb_w * P.mb_h;
    const int ppr = (P.mb_w + NM - 1) / NM;              // items per MB row
    // P.pair_row (NM = 4 only): the stripe's rows from pair_row on are dealt out as items of 2 MBs, so that the launch
    // is whole rounds of 4-MB items plus a short tail of pairs (see d_n_items_stripe)
    const int ppr2 = (P.mb_w + 1) / 2;
    const int row_split = (NM == 4 && P.pair_row > 0 && !P.mb_list) ? min(max(P.pair_row, P.mb_row_begin), P.mb_row_end) : P.mb_row_end;
    const int n_it4 = (row_split - P.mb_row_begin) * ppr;
    const int n_it_stripe = P.mb_list ? P.n_list : n_it4 + (P.mb_row_end - row_split) * ppr2;   // a list: one MB per item
    const int n_items = n_it_stripe * P.num_refs;
    constexpr int NPB = PER_BLOCK ? JMME_NBLK : 1;

    const int bonus_base = P.rdopt ? 0 : d_weighted_cost(P.lambda_factor, 16);
    const unsigned bias = (unsigned)bonus_base;          // keeps (cost + bias) >= 0
    const bool pretest = (!P.rdopt) && P.search_mode == JMME_SEARCH_FASTFULL;
    int patched = -1;

    // tasks: 16 consecutive columns of one run per warp; residual columns gathered several runs per warp.
    // Main tasks come from a small table (no division in the task loop); the residual mapping of a lane
    // is a kernel constant.
    const int nruns = (ncols + K - 1) / K;
    const int nseg = ncols >> 4;
    const int wr = ncols - 16 * nseg;                    // 1..15 (ncols is odd)
    const int G = 16 / wr;
    const int n_main = nruns * nseg;
    const int n_tasks = n_main + (nruns + G - 1) / G;
    // the constant tables of the CTA (first read behind the barrier that follows the first prefetch): built while the
    // first window is on its way, except in a wavefront step (WP), where they overlap the previous step's kernel
    auto build_tables = [&]() {
        if (!KEYG && !KRTAB) {                           // 8 keys per load: one round trip instead of ncand / threads
            const uint4 *src = (const uint4 *)P.spiral_key;
            for (int i = tid; i < (ncand >> 3); i += NW * 32) ((uint4 *)s_key)[i] = src[i];
            for (int i = (ncand & ~7) + tid; i < ncand; i += NW * 32) s_key[i] = P.spiral_key[i];
        }
        for (int i = tid; i < JMME_NT; i += NW * 32)
            s_T[i] = ((unsigned)d_weighted_cost(P.lambda_factor, i) + bias) << JMME_KEY_BITS;
        for (int i = tid; i < n_main; i += NW * 32) {
            const int run = i / nseg, seg = i - run * nseg;
            s_task[i] = (uint16_t)((min(run * K, ncols - K) << 8) | (16 * seg));
        }
        if constexpr (KRTAB) {                           // the context's table (jmme_api.cu): one copy, 16 bytes per load
            const uint4 *src = (const uint4 *)P.kr0;
            for (int i = tid; i < (ncand >> 2); i += NW * 32) ((uint4 *)s_kr)[i] = src[i];
            for (int i = (ncand & ~3) + tid; i < ncand; i += NW * 32) s_kr[i] = P.kr0[i];
        }
    };
    if constexpr (WP) build_tables();
    int res_g = l16 / wr;                                // residual task: run offset and column of this lane
    int res_x = 16 * nseg + (l16 - res_g * wr);
    if (res_g >= G) { res_g = 0; res_x = 16 * nseg; }    // idle lanes repeat lane 0 (idempotent)

    if constexpr (WP) {                                  // (first used after the barriers inside decode_item)
        if (tid < JMME_NBLK * 4) s_wsrc[tid >> 2][tid & 3] = P.wave_tab->src[tid >> 2][tid & 3];
        if (tid < JMME_NBLK) s_wtp[tid] = P.wave_tab->tp[tid];
        // Everything above reads launch constants only and may overlap the tail of the previous step's
        // sub-pel kernel; the committed field is read below.  (No early trigger here: sub-pel CTAs launched
        // while this kernel holds most SMs would crowd onto the few free ones — measured slower.)
        pdl_wait();
    }
    auto decode_item = [&](int item, Item &it) {           // called by all threads of the CTA together
        it.ref = item / n_it_stripe;
        const int idx = item - it.ref * n_it_stripe;
        if (P.mb_list) {
            it.mb = P.mb_list[idx];
            it.mby = it.mb / P.mb_w;
            it.mbx = it.mb - it.mby * P.mb_w;
            it.nmb = 1;
        } else {
            if (idx < n_it4) {
                it.mby = P.mb_row_begin + idx / ppr;
                it.mbx = (idx % ppr) * NM;
                it.nmb = min(NM, P.mb_w - it.mbx);
            } else {                                     // the tail rows in pairs
                const int i2 = idx - n_it4;
                it.mby = row_split + i2 / ppr2;
                it.mbx = (i2 % ppr2) * 2;
                it.nmb = min(2, P.mb_w - it.mbx);
            }
            it.mb = it.mby * P.mb_w + it.mbx;
        }
        const int16_t *pr = P.pred ? P.pred + ((size_t)it.ref * n_mb + it.mb) * NPB * 2 : nullptr;
        if constexpr (WP) {
            const int ks = P.slice_rows ? P.slice_rows : P.mb_h;
            if (tid < 10) {
                int x, y;
                wave_slot_xy(it.mbx, it.mby, tid, x, y);
                s_wnb[tid] = wave_load_nb(P.field_mv, P.field_ref, 4 * P.mb_w, 4 * P.mb_h, 4 * ((it.mby / ks) * ks), x, y);
            }
            __syncthreads();
            int px = 0, py = 0;
            if (tid == 0) {
                wave_predict_block(s_wnb, s_wsrc[0], s_wtp[0], it.ref, 0, 0, px, py);
                s_wpred[0] = (int16_t)px; s_wpred[1] = (int16_t)py;
            }
            __syncthreads();
            if (tid >= 1 && tid < JMME_NBLK) {
                wave_predict_block(s_wnb, s_wsrc[tid], s_wtp[tid], it.ref, s_wpred[0], s_wpred[1], px, py);
                s_wpred[2 * tid] = (int16_t)px; s_wpred[2 * tid + 1] = (int16_t)py;
            }
            __syncthreads();
            if (crank == 0 && tid < JMME_NBLK)             // for the sub-pel kernel and jmme_get_predictors
                *(uint32_t *)(P.pred + (((size_t)it.ref * n_mb + it.mb) * JMME_NBLK + tid) * 2) =
                    (uint32_t)(uint16_t)s_wpred[2 * tid] | ((uint32_t)(uint16_t)s_wpred[2 * tid + 1] << 16);
            pr = s_wpred;
        }
        const int p16x = pr ? d_pred(pr[0]) : 0, p16y = pr ? d_pred(pr[1]) : 0;
        it.cx = d_clamp(p16x / 4, -P.cmax, P.cmax);
        it.cy = d_clamp(p16y / 4, -P.cmax, P.cmax);
    };
    auto prefetch = [&](const Item &it, int buf) {
        const uint8_t *plane = P.planes[it.ref];
        const int gx0 = P.pad + 16 * it.mbx + it.cx - R, gy0 = P.pad + 16 * it.mby + it.cy - R;
        const uint8_t *g = plane + (size_t)gy0 * P.pstride + (gx0 & ~15);
        const int nch = RAWW >> 2;                       // 16-byte chunks per raw row
        if (nch <= 8) {                                  // thread = (row, chunk slot): no division
            const int c = tid & 7;
            if (c < nch)
                for (int row = tid >> 3; row < rows; row += NW * 4)
                    cp_async16(s_raw + row * RAWW + 4 * c, g + (size_t)row * P.pstride + 16 * c);
        } else {
            const int c = tid & 15;
            if (c < nch)
                for (int row = tid >> 4; row < rows; row += NW * 2)
                    cp_async16(s_raw + row * RAWW + 4 * c, g + (size_t)row * P.pstride + 16 * c);
        }
        if (tid < 16 * it.nmb) {                         // current MB(s): 16 rows of 16 bytes each
            const int m = tid >> 4, row = tid & 15;
            cp_async16(s_cur2 + buf * CURW + 64 * m + 4 * row,
                       P.cur + (size_t)min(16 * it.mby + row, P.cur_h - 1) * P.cur_stride + 16 * (it.mbx + m));
        }
        cp_async_commit();
    };
    auto expand = [&](const Item &it) {
        const int gx0 = P.pad + 16 * it.mbx + it.cx - R;
        const int t16 = gx0 & 15, np = RS >> 1;          // word pairs per row (RS is even)
        // (warp-uniform trip count: see the note on loops in front of a barrier in jmme_dev.cuh)
        for (int xp0 = 0; xp0 < np; xp0 += 32) {         // this lane's word pair(s): fixed across rows
            const int xp = xp0 + lane;
            if (xp >= np) continue;
            const int o0 = t16 + 2 * xp, i0 = o0 >> 2, b0 = o0 & 3;
            // words at byte offsets o0 and o0+1 out of three aligned raw words (byte permute)
            const unsigned sel0 = 0x3210u + 0x1111u * b0;
            const unsigned sel1 = b0 == 3 ? 0x3210u : 0x3210u + 0x1111u * (b0 + 1);
#pragma unroll 4
            for (int row = warp; row < rows; row += NW) {
                const uint32_t *raw = s_raw + row * RAWW + i0;
                const uint32_t a0 = raw[0], a1 = raw[1], a2 = raw[2];
                uint2 v;
                v.x = __byte_perm(a0, a1, sel0);
                v.y = b0 == 3 ? __byte_perm(a1, a2, sel1) : __byte_perm(a0, a1, sel1);
                *(uint2 *)(s_win + row * RS + 2 * xp) = v;
            }
        }
        for (int i = tid; i < 48 * NM; i += NW * 32) s_best[i] = 0xFFFFFFFFu;
        const int16_t *pr = WP ? s_wpred : (P.pred ? P.pred + ((size_t)it.ref * n_mb + it.mb) * NPB * 2 : nullptr);
        if constexpr (!PER_BLOCK) {
            for (int i = tid; i < ncols; i += NW * 32) {
                const int px = pr ? d_pred(pr[0]) : 0, py = pr ? d_pred(pr[1]) : 0;
                s_bx[i] = (uint8_t)d_se_bits(4 * (it.cx + i - R) - px);
                s_by[i] = (uint8_t)d_se_bits(4 * (it.cy + i - R) - py);
            }
        } else {
            // [half][offset][local block] bits * 4.  thread = (word q of a row = 4 local blocks, row slot): the
            // eight predictors it needs (4 blocks x 2 halves) are loaded once, no division in the loop
            constexpr int TPR = NW * 32 / 6;                            // rows per pass
            const int q = tid % 6, r0 = tid / 6;
            if (r0 < TPR) {
                int px[2][4], py[2][4];
#pragma unroll
                for (int hf = 0; hf < 2; hf++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int b = 4 * q + j, gb = b < NL ? kGT[b] + hf * kDL[b] : 0;
                        px[hf][j] = d_pred(pr[2 * gb]); py[hf][j] = d_pred(pr[2 * gb + 1]);
                    }
                for (int row = r0; row < 2 * ncols; row += TPR) {
                    const int hf = row >= ncols, o = row - (hf ? ncols : 0);
                    const int vx = 4 * (it.cx + o - R), vy = 4 * (it.cy + o - R);
                    uint32_t wx = 0, wy = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        wx |= (uint32_t)(4 * d_se_bits(vx - (hf ? px[1][j] : px[0][j]))) << (8 * j);
                        wy |= (uint32_t)(4 * d_se_bits(vy - (hf ? py[1][j] : py[0][j]))) << (8 * j);
                    }
                    ((uint32_t *)s_bx)[row * 6 + q] = wx;
                    ((uint32_t *)s_by)[row * 6 + q] = wy;
                }
            }
        }
        const int idx00 = (R - it.cy) * ncols + (R - it.cx);
        if (!KEYG && !KRTAB && pretest && tid == 0) {    // "(0,0) first": key 0 wins every tie
            if (patched >= 0 && patched != idx00) s_key[patched] = P.spiral_key[patched];
            s_key[idx00] = 0;
        }
        if (pretest) patched = idx00;
    };

    Item cur_it, nxt_it;
    int item = blockIdx.x / CL, buf = 0;
    int item_stride = gridDim.x / CL, item_end = n_items;
    // the sub-pel kernel may start as soon as SMs are free: it waits per MB (P.ready), not for this kernel's end
    if (P.ready) pdl_trigger();
    const int TPI = n_tasks * NM;                        // tasks of an item
    int g0 = 0, g1 = 0;                                  // BAL: this CTA's range of the stripe's tasks
    if constexpr (BAL) {
        const long long gt = (long long)n_items * TPI;
        g0 = (int)(gt * blockIdx.x / gridDim.x);
        g1 = (int)(gt * (blockIdx.x + 1) / gridDim.x);
        if (g0 >= g1) return;
        item = g0 / TPI; item_end = (g1 - 1) / TPI + 1; item_stride = 1;
    }
    if (item >= item_end) return;                        // (the same for every CTA of a cluster)
    decode_item(item, cur_it);
    prefetch(cur_it, 0);
    if constexpr (!WP) build_tables();
    cp_async_wait_all();
    __syncthreads();
    expand(cur_it);
    __syncthreads();

    for (; item < item_end; item += item_stride) {
        const int nxt = item + item_stride;
        const bool has_next = nxt < item_end;
        if (has_next) {
            decode_item(nxt, nxt_it);
            prefetch(nxt_it, buf ^ 1);
        }
        const int cx = cur_it.cx, cy = cur_it.cy;
        const int bonus = (cur_it.ref == 0) ? bonus_base : 0;
        const int x00 = R - cx, y00 = R - cy;
      for (int m = 0; m < cur_it.nmb; m++) {               // the MB(s) of this item, one after the other
        int lo = 0, hi = n_tasks;                        // this CTA's tasks of the MB
        if constexpr (BAL) {
            lo = max(g0 - item * TPI - n_tasks * m, 0);
            hi = min(g1 - item * TPI - n_tasks * m, n_tasks);
            if (lo >= hi) continue;                      // (uniform over the CTA)
        }
        const uint32_t *s_cur = s_cur2 + buf * CURW + 64 * m;
        const int wx = 16 * m;                           // window column of this MB's offset 0
        uint32_t *s_bestm = s_best + 48 * m;

        // The current MB is the same for every lane; an opaque zero lane offset keeps ptxas from
        // parking it in uniform registers (it then pays a UR->R move per VABSDIFF4 operand).
        unsigned lz;
        asm volatile("and.b32 %0, %1, 0;" : "=r"(lz) : "r"(lane));
        uint32_t cur[8][4];                              // this half's 8 rows of the current MB
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint4 v = *(const uint4 *)(s_cur + lz + 4 * (8 * half + r));
            cur[r][0] = v.x; cur[r][1] = v.y; cur[r][2] = v.z; cur[r][3] = v.w;
        }

        if (bonus != 0 && warp == 0 && crank == 0 && lo == 0) {     // 16x16 at MV (0,0) with its bonus (BAL: by the CTA that has task 0)
            unsigned s = 0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int i = lane + 32 * h, row = i >> 2, j = i & 3;
                s = sad4(s_cur[i], s_win[(y00 + row) * RS + x00 + wx + 4 * j], s);
            }
            s = __reduce_add_sync(0xFFFFFFFFu, s);
            if (lane == 0) {
                const unsigned k00 = pretest ? 0u : (unsigned)P.spiral_key[y00 * ncols + x00];
                const unsigned bits00 = PER_BLOCK ? (s_bx[x00 * 24 + 19] + s_by[y00 * 24 + 19]) >> 2    // local 19 = 16x16
                                                  : s_bx[x00] + s_by[y00];
                const unsigned v = (s << JMME_KEY_BITS) + s_T[bits00] + k00 -
                                   ((unsigned)bonus << JMME_KEY_BITS);
                atomicMin(&s_bestm[0], v);
            }
        }

        uint32_t best[NL];
#pragma unroll
        for (int b = 0; b < NL; b++) best[b] = 0xFFFFFFFFu;

        // second MB: the warp that starts at task 0 (one task more than the others) rotates
        const int t0 = lo + (warp + NM * NW - m) % NW + (int)crank * NW;
        constexpr int TS = NW * CL;                              // task stride
        unsigned e_next = t0 < n_main ? s_task[t0] : 0u;         // task descriptor, fetched one task ahead
        for (int task = t0; task < hi; task += TS) {
            int ybase, xoff;
            if (task < n_main) {
                ybase = e_next >> 8;
                xoff = (e_next & 255) + l16;
            } else {
                ybase = min(min((task - n_main) * G + res_g, nruns - 1) * K, ncols - K);
                xoff = res_x;
            }
            if (task + TS < n_main) e_next = s_task[task + TS];
            const uint32_t *base = s_win + (ybase + 8 * half) * RS + xoff + wx;

            unsigned acc[K][8];
#pragma unroll
            for (int k = 0; k < K; k++)
#pragma unroll
                for (int i = 0; i < 8; i++) acc[k][i] = 0;

            unsigned bx0 = 0;
            uint32_t bxw[6];                             // PER_BLOCK: bits*4 of this column for the 22 local blocks
            if constexpr (PER_BLOCK) {
                const uint2 *q = (const uint2 *)(s_bx + (half * ncols + xoff) * 24);
#pragma unroll
                for (int i = 0; i < 3; i++) { const uint2 v = q[i]; bxw[2 * i] = v.x; bxw[2 * i + 1] = v.y; }
            } else if constexpr (!KRTAB) {
                bx0 = s_bx[xoff];
            }
            auto pack = [&](int k, unsigned (&pk)[NL]) {
                const int yoff = ybase + k;
                unsigned key = 0;
                if constexpr (KRTAB) {
                } else if constexpr (KEYG) {
                    const int ki = yoff * ncols + xoff;
                    key = (ki == patched) ? 0u : (unsigned)__ldg(P.spiral_key + ki);    // patched = MV (0,0) pre-test
                } else {
                    key = s_key[yoff * ncols + xoff];
                }
                unsigned o[NL];
                const unsigned(&s)[8] = acc[k];
#pragma unroll
                for (int i = 0; i < 8; i++) o[11 + i] = s[i];                         // 4x4
#pragma unroll
                for (int i = 0; i < 4; i++) o[7 + i] = s[i] + s[4 + i];               // 4x8
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int i = 0; i < 2; i++) o[3 + 2 * j + i] = s[4 * j + 2 * i] + s[4 * j + 2 * i + 1];   // 8x4
                o[1] = o[3] + o[5];                                                   // 8x8 left, right
                o[2] = o[4] + o[6];
                o[0] = o[1] + o[2];                                                   // 16x8
                const unsigned pl = __shfl_xor_sync(0xFFFFFFFFu, o[1], 16);          // the other half's 8x8 sums
                const unsigned pr8 = __shfl_xor_sync(0xFFFFFFFFu, o[2], 16);
                o[20] = o[1] + pl;                                                    // 8x16 left
                o[21] = o[2] + pr8;                                                   // 8x16 right
                o[19] = o[20] + o[21];                                                // 16x16
                if constexpr (KRTAB) {
                    const unsigned kr = s_kr[yoff * ncols + xoff];
#pragma unroll
                    for (int b = 0; b < NL; b++) pk[b] = (o[b] << JMME_KEY_BITS) + kr;
                } else if constexpr (!PER_BLOCK) {
                    const unsigned kr = s_T[bx0 + s_by[yoff]] + key;
#pragma unroll
                    for (int b = 0; b < NL; b++) pk[b] = (o[b] << JMME_KEY_BITS) + kr;
                } else {
                    // byte-wise bits*4 sums of four blocks per add; each byte is the offset of its rate in s_T
                    const uint2 *q = (const uint2 *)(s_by + (half * ncols + yoff) * 24);
                    uint32_t sw[6];
#pragma unroll
                    for (int i = 0; i < 3; i++) { const uint2 v = q[i]; sw[2 * i] = v.x + bxw[2 * i]; sw[2 * i + 1] = v.y + bxw[2 * i + 1]; }
                    if constexpr (LIN) {
                        const unsigned lam8k = (unsigned)(P.lambda_factor >> 16) << 13;
                        const unsigned keyb = key + (bias << JMME_KEY_BITS);
#pragma unroll
                        for (int b = 0; b < NL; b++) {
                            const unsigned b4 = __byte_perm(sw[b >> 2], 0, 0x4440 | (b & 3));
                            pk[b] = (o[b] << JMME_KEY_BITS) + (b4 * lam8k + keyb);
                        }
                    } else {
#pragma unroll
                        for (int b = 0; b < NL; b++) {
                            const unsigned off = __byte_perm(sw[b >> 2], 0, 0x4440 | (b & 3));
                            pk[b] = (o[b] << JMME_KEY_BITS) + *(const uint32_t *)((const char *)s_T + off) + key;
                        }
                    }
                }
            };

            // Row loop.  Candidate k is complete after row 7+k; its partition sums, packs (FMA pipe) and
            // minima are emitted right there, between the VABSDIFF4 (ALU pipe) of the later candidates,
            // two candidates at a time so that min(best, min(a, b)) is one VIMNMX3.
#pragma unroll
            for (int rr = 0; rr < 8 + K - 1; rr++) {
                const uint32_t *rp = base + rr * RS;
                const unsigned r0 = rp[0], r1 = rp[4], r2 = rp[8], r3 = rp[12];
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int cr = rr - k;
                    if (cr >= 0 && cr < 8) {
                        const int a = (cr >> 2) * 4;
                        acc[k][a + 0] = sad4(cur[cr][0], r0, acc[k][a + 0]);
                        acc[k][a + 1] = sad4(cur[cr][1], r1, acc[k][a + 1]);
                        acc[k][a + 2] = sad4(cur[cr][2], r2, acc[k][a + 2]);
                        acc[k][a + 3] = sad4(cur[cr][3], r3, acc[k][a + 3]);
                    }
                }
                const int kc = rr - 7;                   // candidate completed by this row
                if (kc >= 1 && (kc & 1)) {
                    unsigned pa[NL], pb[NL];
                    pack(kc - 1, pa);
                    pack(kc, pb);
#pragma unroll
                    for (int b = 0; b < NL; b++) best[b] = min(best[b], min(pa[b], pb[b]));
                } else if (kc == K - 1 && !(kc & 1)) {
                    unsigned pa[NL];
                    pack(kc, pa);
#pragma unroll
                    for (int b = 0; b < NL; b++) best[b] = min(best[b], pa[b]);
                }
            }
        }

        // ---- reduce: per half over its 16 lanes (the other half contributes the identity), lane b
        //      keeps global block b, two shared atomicMin per warp ---------------------------------
        {
            unsigned m0 = 0xFFFFFFFFu, m1 = 0xFFFFFFFFu;
            auto keep = [&](int gb, unsigned m) {
                if (gb < 32) m0 = (lane == gb) ? m : m0;
                else m1 = (lane == gb - 32) ? m : m1;
            };
#pragma unroll
            for (int b = 0; b < NL; b++) {
                if (kDL[b] == 0) {
                    keep(kGT[b], __reduce_min_sync(0xFFFFFFFFu, best[b]));
                } else {
                    keep(kGT[b], __reduce_min_sync(0xFFFFFFFFu, half == 0 ? best[b] : 0xFFFFFFFFu));
                    keep(kGT[b] + kDL[b], __reduce_min_sync(0xFFFFFFFFu, half == 1 ? best[b] : 0xFFFFFFFFu));
                }
            }
            atomicMin(&s_bestm[lane], m0);
            if (lane < JMME_NBLK - 32) atomicMin(&s_bestm[32 + lane], m1);
        }
      }
        cp_async_wait_all();
        __syncthreads();
        if constexpr (CL > 1) {                          // partial minima of the other CTAs -> CTA 0
            auto cluster = cooperative_groups::this_cluster();
            cluster.sync();
            if (crank != 0 && tid < JMME_NBLK) atomicMin(cluster.map_shared_rank(&s_best[tid], 0), s_best[tid]);
            cluster.sync();
        }
        if constexpr (BAL) {                             // packed minima -> global (combined across CTAs, decoded by the consumer)
            for (int i = tid; i < JMME_NBLK * cur_it.nmb; i += NW * 32) {
                const int m = i / JMME_NBLK, b = i - JMME_NBLK * m;
                const unsigned v = s_best[48 * m + b];
                if (v != 0xFFFFFFFFu) atomicMin(P.gbest + ((size_t)cur_it.ref * n_mb + cur_it.mb + m) * JMME_NBLK + b, v);
            }
        } else if (crank == 0)
        for (int i = tid; i < JMME_NBLK * cur_it.nmb; i += NW * 32) {
            const int m = i / JMME_NBLK, b = i - JMME_NBLK * m;
            const unsigned v = s_best[48 * m + b];
            const unsigned key = v & JMME_KEY_MASK;
            int mvx = 0, mvy = 0;
            if (key) {
                d_spiral_xy((int)key - 1, mvx, mvy);
                mvx += cx; mvy += cy;
            }
            BlkRes r;
            r.mvx = (int16_t)(4 * mvx);
            r.mvy = (int16_t)(4 * mvy);
            r.cost = (int)(v >> JMME_KEY_BITS) - (int)bias;
            P.res[((size_t)cur_it.ref * n_mb + cur_it.mb + m) * JMME_NBLK + b] = r;
        }
        if (P.ready) {                                   // the integer result of this item's MBs is complete: raise their flags
            __threadfence();
            __syncthreads();
            if (crank == 0 && tid < cur_it.nmb) {
                __threadfence();
                *(volatile int *)(P.ready + (size_t)cur_it.ref * n_mb + cur_it.mb + tid) = 1;
            }
        }
        if (!has_next) break;
        __syncthreads();
        expand(nxt_it);
        __syncthreads();
        cur_it = nxt_it;
        buf ^= 1;
    }
}

template <int K, int NW, int MINB, bool PER_BLOCK, int RS_CT, bool KEYG, bool KRTAB = false, int NMB = 1, int CL = 1,
          bool WP = false, bool LIN = false, bool BAL = false>
cudaError_t launch_tb(const SearchParams &P, int num_sms, cudaStream_t st)
{
    TbLayout L(P.R, PER_BLOCK, !KEYG && !KRTAB, K, KRTAB, NMB);
    size_t bytes = (size_t)L.total_words * 4;
    auto kern = me_int_tb_kernel<K, NW, MINB, PER_BLOCK, RS_CT, KEYG, KRTAB, NMB, CL, WP, LIN, BAL>;
    static KernelState ks;                               // shared-memory opt-in and occupancy, per device
    int c_occ = 0;
    cudaError_t e = jmme_kernel_occupancy(kern, ks, NW * 32, bytes, &c_occ);
    if (e != cudaSuccess) return e;
    snprintf(jmme_kernel_name_buf(), JMME_KNAME_LEN,
             "me_int_tb_kernel<K=%d,NW=%d,MINB=%d,PER_BLOCK=%d,RS_CT=%d,KEYG=%d,KRTAB=%d,NMB=%d,CL=%d,WP=%d,LIN=%d,BAL=%d>", K, NW, MINB,
             (int)PER_BLOCK, RS_CT, (int)KEYG, (int)KRTAB, NMB, CL, (int)WP, (int)LIN, (int)BAL);
    int n_items = (P.mb_list ? P.n_list : (P.mb_row_end - P.mb_row_begin) * ((P.mb_w + NMB - 1) / NMB)) * P.num_refs;
    if (NMB == 4 && P.pair_row > 0 && !P.mb_list) {
        const int rs = std::min(std::max(P.pair_row, P.mb_row_begin), P.mb_row_end);
        n_items = ((rs - P.mb_row_begin) * ((P.mb_w + 3) / 4) + (P.mb_row_end - rs) * ((P.mb_w + 1) / 2)) * P.num_refs;
    }
    if (n_items <= 0) return cudaSuccess;
    if (BAL) {
        // every resident CTA slot gets an equal share of the tasks; at least 8 tasks per CTA (a tiny stripe must not
        // pay one window staging per task)
        const int nruns = (P.ncols + K - 1) / K, wr = P.ncols & 15, n_tasks = nruns * (P.ncols >> 4) + (nruns + 16 / wr - 1) / (16 / wr);
        const long long gt = (long long)n_items * n_tasks * NMB;
        const int grid = (int)std::max<long long>(1, std::min<long long>(num_sms * c_occ, gt / 8));
        kern<<<grid, NW * 32, bytes, st>>>(P);
        return cudaGetLastError();
    }
    if (CL > 1 || P.pdl) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[2];
        int na = 0;
        if (CL > 1) {
            at[na].id = cudaLaunchAttributeClusterDimension;
            at[na].val.clusterDim.x = CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
            na++;
        }
        if (P.pdl) {
            at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[na].val.programmaticStreamSerializationAllowed = 1;
            na++;
        }
        cfg.gridDim = dim3((unsigned)(CL * min(n_items, max(num_sms * c_occ / CL, 1))));
        cfg.blockDim = dim3(NW * 32);
        cfg.dynamicSmemBytes = bytes; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = na;
        return cudaLaunchKernelEx(&cfg, kern, P);
    }
    int grid = min(n_items, num_sms * c_occ);
    kern<<<grid, NW * 32, bytes, st>>>(P);
    return cudaGetLastError();
}

// balanced-range launch of a shape that has one (else the launch is refused: jmme_me_int_balanced never asks for it)
template <int K, int NW, int MINB, int RS, int NMB, bool OK>
cudaError_t launch_bal(const SearchParams &P, int num_sms, cudaStream_t st)
{
    if constexpr (OK) return launch_tb<K, NW, MINB, false, RS, false, true, NMB, 1, false, false, true>(P, num_sms, st);
    else return cudaErrorInvalidValue;
}

}  // namespace

// shape 7: 4 warps, >= 4 CTAs/SM (<= 128 registers)   shape 8: 4 warps, >= 3 CTAs/SM (<= 168 registers)
// shape 5: 6 warps, >= 2 CTAs/SM; shape 4: 12 warps, 1 CTA/SM — fewer, faster items for short stripes
// shape 6: 3 warps, >= 4 CTAs/SM (<= 168 registers): 45 tasks per MB at R = 32, K = 6 split 15/15/15
// shape 9: 6 warps, >= 2 CTAs/SM (<= 168 registers), spiral keys read from global memory: the shape for
//          R > 32, where the window (80 KB at R = 64) leaves room for only two CTAs per SM
cudaError_t jmme_launch_me_int_tb(const SearchParams &P, int num_sms, int K, int shape, cudaStream_t st)
{
    const bool pb = P.pred_policy == JMME_PRED_PER_BLOCK;
    const bool lin = (P.lambda_factor & 0xFFFF) == 0 && P.tune_lin;       // integer lambda: linear rate (LIN)
    if (pb && lin && !P.mb_list && K == 6 && shape == 8)                  // the default per-block kernel
        return launch_tb<6, 4, 3, true, 0, false, false, 1, 1, false, true>(P, num_sms, st);
    if (pb && P.mb_list && K == 6 && shape == 4) {       // a wavefront step: spread each MB over a cluster
        const int n_items = P.n_list * P.num_refs;
        const int cmax = P.tune_cluster;                     // 1 = no clusters
        if (P.wave_tab) {                                    // predictors computed in the kernel
            const int cl = (cmax >= 4 && 4 * n_items <= num_sms) ? 4 : ((cmax >= 2 && 2 * n_items <= num_sms) ? 2 : 1);
            if (lin) {
                if (cl == 4) return launch_tb<6, 12, 1, true, 0, false, false, 1, 4, true, true>(P, num_sms, st);
                if (cl == 2) return launch_tb<6, 12, 1, true, 0, false, false, 1, 2, true, true>(P, num_sms, st);
                return launch_tb<6, 12, 1, true, 0, false, false, 1, 1, true, true>(P, num_sms, st);
            }
            if (cl == 4) return launch_tb<6, 12, 1, true, 0, false, false, 1, 4, true>(P, num_sms, st);
            if (cl == 2) return launch_tb<6, 12, 1, true, 0, false, false, 1, 2, true>(P, num_sms, st);
            return launch_tb<6, 12, 1, true, 0, false, false, 1, 1, true>(P, num_sms, st);
        }
        if (cmax >= 4 && 4 * n_items <= num_sms) return launch_tb<6, 12, 1, true, 0, false, false, 1, 4>(P, num_sms, st);
        if (cmax >= 2 && 2 * n_items <= num_sms) return launch_tb<6, 12, 1, true, 0, false, false, 1, 2>(P, num_sms, st);
    }
    if (P.wave_tab) return cudaErrorInvalidValue;            // the host asks for in-kernel prediction only with this shape
#define TB(KK, SH, NWW, MB, KG, BALOK)                                                          \
    if (K == KK && shape == SH) {                                                          \
        if (pb) return launch_tb<KK, NWW, MB, true, 0, KG>(P, num_sms, st);                \
        if (P.R == 32 && !P.pred && !KG) {                                                                     \
            const int grp = P.tune_group;                           /* MBs per item, default 2 */             \
            if (P.int_packed) {                                     /* balanced task ranges (BAL) */           \
                if (grp >= 4) return launch_bal<KK, NWW, MB, 126, 4, BALOK>(P, num_sms, st);                    \
                if (grp >= 2) return launch_bal<KK, NWW, MB, 94, 2, BALOK>(P, num_sms, st);                     \
                return launch_bal<KK, NWW, MB, 78, 1, BALOK>(P, num_sms, st);                                   \
            }                                                                                                   \
            if (grp >= 4) return launch_tb<KK, NWW, MB, false, 126, false, true, 4>(P, num_sms, st);            \
            if (grp >= 2) return launch_tb<KK, NWW, MB, false, 94, false, true, 2>(P, num_sms, st);             \
            return launch_tb<KK, NWW, MB, false, 78, false, true, 1>(P, num_sms, st);                           \
        }                                                                                                       \
        if (P.R == 32) return launch_tb<KK, NWW, MB, false, 78, KG>(P, num_sms, st);       \
        if (P.R == 64) return launch_tb<KK, NWW, MB, false, 142, KG>(P, num_sms, st);      \
        return launch_tb<KK, NWW, MB, false, 0, KG>(P, num_sms, st);                       \
    }
    // (BALOK: the shapes that have balanced-range instantiations — the default and the two the tests sweep)
    TB(4, 7, 4, 4, false, false)
    TB(4, 8, 4, 3, false, true) TB(6, 8, 4, 3, false, true) TB(8, 8, 4, 3, false, false)
    TB(6, 9, 6, 2, true, false)
    TB(6, 6, 3, 4, false, false)
    TB(6, 5, 6, 2, false, true) TB(6, 4, 12, 1, false, false)
#undef TB
    return cudaErrorInvalidValue;
}

// Does the integer search of P run with balanced task ranges and leave its result as packed minima in P.gbest
// (SearchParams::int_packed)?  Mirrors the dispatch above and in jmme_launch_me_int: the default two-thread kernel,
// zero predictors, R = 32, whole stripes.
bool jmme_me_int_balanced(const SearchParams &P, int variant, int num_sms, bool forced)
{
    if (!P.tune_split || !P.gbest || !P.kr0) return false;
    // measured (tools/sweep_split.py, 1080p, search + sub-pel): whole items with the early sub-pel start beat the
    // balanced ranges at every stripe size (68 rows 0.412 vs 0.433 ms, 17 rows 0.111 vs 0.125 ms), so balance is
    // opt-in (jmme_tuning.balance = 1)
    (void)num_sms;
    if (!forced) return false;
    if (P.metric[0] == JMME_DIST_SSE || P.cost_domain || P.R != 32 || P.pred || P.mb_list) return false;
    if (P.blocktype_mask == JMME_MASK_16x16) return false;
    if (variant <= 0) variant = 68;
    return variant == 68 || variant == 65 || variant == 48;      // the shapes with BAL instantiations (TB(..., true))
}

// Does the integer search of P go to a me_int_tb_kernel launch that raises the per-MB ready flags (SearchParams::ready:
// the sub-pel kernel then starts early)?  Mirrors jmme_launch_me_int: the two-thread kernel on a whole stripe (no MB
// list, no balanced ranges: an MB has one writer).
bool jmme_me_int_raises_flags(const SearchParams &P, int variant)
{
    if (P.int_packed || P.mb_list || P.metric[0] == JMME_DIST_SSE || P.cost_domain) return false;
    if (P.blocktype_mask == JMME_MASK_16x16) return false;
    if (variant <= 0) variant = P.R <= 32 ? 68 : 51;
    const int K = variant / 10, c = variant % 10;
    return c >= 4 && K <= P.ncols;
}
