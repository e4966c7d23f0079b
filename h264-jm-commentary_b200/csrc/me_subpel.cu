// me_subpel.cu — half- then quarter-pel refinement of all 41 blocks of a macroblock, and the
// final per-block choice of the reference picture.
//
// Stands in for JM's SubPelBlockMotionSearch with SATD ‖ HadamardSAD4x4
// (a11) and the REF_COST comparison of PartitionMotionSearch's caller (a4).  Conventions: DESIGN.md §2.
//
// Mapping: one CTA per (reference, macroblock), thread = (blocktype, 4x4 cell); see me_subpel_kernel.
#include <algorithm>

#include "jmme_dev.cuh"

namespace {

__constant__ int8_t c_sp9[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};
// the same table as compile-time constants for the unrolled position loop
__device__ constexpr int c_sp9h[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};
// block index of the 4x4 cell (cx4, cy4) for each blocktype
__device__ __forceinline__ int block_of_cell(int t, int cx4, int cy4)
{
    switch (t) {
    case 1: return 0;
    case 2: return 1 + (cy4 >> 1);
    case 3: return 3 + (cx4 >> 1);
    case 4: return 5 + 2 * (cy4 >> 1) + (cx4 >> 1);
    case 5: return 9 + 2 * cy4 + (cx4 >> 1);
    case 6: return 17 + 4 * (cy4 >> 1) + cx4;
    default: return 25 + 4 * cy4 + cx4;
    }
}

// need[t]: XOR offsets (within the 16 cells of an MB, cell = 4*cy4 + cx4) that gather the cells of one
// block of blocktype t: 16x16 all, 16x8 {1,2,4}, 8x16 {1,4,8}, 8x8 {1,4}, 8x4 {1}, 4x8 {4}, 4x4 none
__device__ __forceinline__ int need_of_type(int t) { return (0x0415D7F0u >> (4 * t)) & 15; }

// One CTA (4 warps) per (reference, macroblock).  Thread = (blocktype, 4x4 cell): half-warp h of warp w
// owns blocktype 2w+h+1 (the last half-warp is idle), lane&15 is the cell.  For each candidate
// position the thread computes its cell's SATD (or SAD) from four unaligned 32-bit reference reads,
// the cells of a block are summed with a masked XOR-shuffle butterfly, and every lane of the block
// runs the same strict-< scan over the positions — no shared memory, no barriers, no atomics.
// LAT (the MB lists of the in-frame median wavefront: a few dozen CTAs on the whole GPU, nothing to hide the
// L2 latency behind): the plane words of all nine positions of a step are requested before the first is used.
// NG = 3 (with LAT): three groups of 128 threads share the positions of a step (3 + 3 + 3, then 2 + 3 + 3),
// each runs the strict-< scan over its own and the groups' (cost, position) minima meet in shared memory:
// the lowest position among equal costs wins, as in the sequential scan.  A third of the dependent chain.
// EXT (P.ext): the general form of the stages — either cost domain, SAD / SSE / Hadamard per stage, 8x8
// Hadamard for blocktypes 1..4 (two cross-cell butterfly stages over SHFL before the 4x4 transform: the four
// cells of an 8x8 are lanes c, c^1, c^4, c^5), chroma ME (2x2 chroma samples per cell and plane, eighth-pel
// bilinear on the fly).  The legacy instantiation keeps its instruction stream.
template <bool LAT, int NG, bool EXT>
__global__ void __launch_bounds__(128 * NG) me_subpel_kernel(const SearchParams P)
{
    __shared__ long long s_pk[NG > 1 ? 2 : 1][NG > 1 ? NG : 1][NG > 1 ? JMME_NBLK : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = (tid >> 5) & 3, grp = tid >> 7;
    const int t = 2 * warp + (lane >> 4) + 1;            // blocktype 1..7 (8 = idle)
    const int cell = lane & 15, cx4 = cell & 3, cy4 = cell >> 2;
    const int n_mb_stripe = d_n_units(P);
    const int n_mb = P.mb_w * P.mb_h;
    const int item = blockIdx.x;
    const int ref = item / n_mb_stripe;
    const int mb = d_unit_mb(P, item - ref * n_mb_stripe);
    const int mby = mb / P.mb_w, mbx = mb - mby * P.mb_w;
    const bool active = t <= 7 && ((P.blocktype_mask >> t) & 1);
    const int b = t <= 7 ? block_of_cell(t, cx4, cy4) : 0;
    const int need = t <= 7 ? need_of_type(t) : 0;
    BlkRes *res = P.res + ((size_t)ref * n_mb + mb) * JMME_NBLK;
    const bool has_bonus = !P.rdopt && ref == 0 && b == 0;
    const size_t psz = (size_t)P.pstride * P.pheight;
    const uint8_t *planes = P.planes[ref];
    const bool had8w = EXT && P.t8 && warp < 2;          // this warp's blocktypes (1..4) use the 8x8 transform

    // this cell of the current MB: raw words (SAD) and 16-bit lane pairs (c0|c2<<16), (c1|c3<<16) per row (SATD)
    uint32_t cw[4], ca[4], cb[4];
    {
        const uint8_t *cp = P.cur + 16 * mbx + 4 * cx4;
#pragma unroll
        for (int y = 0; y < 4; y++) {
            cw[y] = *(const uint32_t *)(cp + (size_t)min(16 * mby + 4 * cy4 + y, P.cur_h - 1) * P.cur_stride);
            ca[y] = __byte_perm(cw[y], 0, 0x4240);
            cb[y] = __byte_perm(cw[y], 0, 0x4341);
        }
    }
    int cc[2][4] = {};                                   // EXT chroma ME: this cell's 2x2 samples of Cb and Cr
    if constexpr (EXT) {
        if (P.chroma_me) {
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const uint8_t *q = P.cur_c[k] + (size_t)(8 * mby + 2 * cy4 + j) * P.cur_cs + 8 * mbx + 2 * cx4;
                    cc[k][2 * j] = q[0]; cc[k][2 * j + 1] = q[1];
                }
        }
    }
    if constexpr (LAT) {                                 // (the loads above read the current picture only)
        pdl_trigger();
        pdl_wait();
    }
    if (P.ready) {
        // early start: this kernel runs beside the last round of the search kernel; wait for this MB's integer result
        // (bounded: a flag that never comes must not hang the GPU — the parity tests would show it), lower the flag for
        // the next search, and read the result past the L1 (a neighbour's line may have been cached before it was final)
        if (tid == 0) {
            volatile int *f = P.ready + (size_t)ref * n_mb + mb;
            unsigned long long t0 = 0, t1 = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while (*f == 0) {
                __nanosleep(200);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 200000000ull) break;           // 0.2 s
            }
            __threadfence();
            *f = 0;
        }
        __syncthreads();
    }
    int mvx, mvy, mn, px = 0, py = 0;
    {
        // (a balanced integer search leaves packed minima: decoded here, reset by the block's owner lane below)
        BlkRes r;
        if (P.int_packed) {
            r = d_unpack_int(P, P.gbest[((size_t)ref * n_mb + mb) * JMME_NBLK + b]);
        } else if (P.ready) {
            const uint2 v = __ldcg((const uint2 *)(res + b));
            r.mvx = (int16_t)(v.x & 0xFFFF); r.mvy = (int16_t)(v.x >> 16); r.cost = (int)v.y;
        } else {
            r = res[b];
        }
        mvx = r.mvx; mvy = r.mvy;
        mn = r.cost;
        if (P.pred) {
            const int npb = P.pred_policy == JMME_PRED_PER_BLOCK ? JMME_NBLK : 1;
            const int16_t *pr = P.pred + ((size_t)ref * n_mb + mb) * npb * 2 + (npb == 1 ? 0 : 2 * b);
            px = d_pred(pr[0]); py = d_pred(pr[1]);
        }
    }
    // reference position of this cell at MV (0,0), in padded-plane coordinates
    const int rx0 = P.pad + 16 * mbx + 4 * cx4, ry0 = P.pad + 16 * mby + 4 * cy4;
    const int pw = P.pstride >> 2;

    int prev_metric = P.metric[0];
    for (int step = 2; step >= 1; step--) {
        // a stage whose metric differs from the previous stage's (or any stage with chroma ME) starts at
        // position 0 with the minimum reset (JM start_me_refinement_hp / _qp)
        const int st = 3 - step, metric = P.metric[st];
        const bool restart = metric != prev_metric || (EXT && P.chroma_me);
        prev_metric = metric;
        if (restart) mn = INT_MAX;
        const int pos0 = restart ? 0 : 1;
        const unsigned lf = (unsigned)P.lf[st];          // legacy: lambda_factor * bits < 2^31, the 32-bit product is exact
        const int bonus = has_bonus ? d_wcost(EXT ? P.cost_domain : 0, P.lf[st], 16) : 0;
        const int ox = mvx, oy = mvy;
        int best = 0;
        // MV bits of the three x and three y offsets of this step (the nine positions combine them)
        int bitx[3], bity[3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            bitx[j] = d_se_bits(ox + (j - 1) * step - px);
            bity[j] = d_se_bits(oy + (j - 1) * step - py);
        }
        uint32_t raw[LAT ? 9 : 1][8];
        if constexpr (LAT) {
#pragma unroll
            for (int pos = 0; pos < 9; pos++) {
                if (pos < pos0 || !active || (NG > 1 && pos / 3 != grp)) continue;
                const int qx = ox + step * c_sp9h[pos][0], qy = oy + step * c_sp9h[pos][1];
                const size_t off = psz * ((qy & 3) * 4 + (qx & 3)) + (size_t)(ry0 + (qy >> 2)) * P.pstride + (rx0 + (qx >> 2));
                const uint32_t *rp = (const uint32_t *)(planes + (off & ~(size_t)3));
#pragma unroll
                for (int y = 0; y < 4; y++) { raw[pos][2 * y] = __ldg(rp + y * pw); raw[pos][2 * y + 1] = __ldg(rp + y * pw + 1); }
            }
        }
#pragma unroll
        for (int pos = 0; pos < 9; pos++) {
            if (pos < pos0 || (NG > 1 && pos / 3 != grp)) continue;  // (uniform per warp)
            const int sx = c_sp9h[pos][0], sy = c_sp9h[pos][1];     // compile-time after unrolling
            const int qx = ox + step * sx, qy = oy + step * sy;
            int v = 0, vc = 0;                                       // luma and (EXT) chroma distortion of this cell
            uint32_t w[4] = {0, 0, 0, 0};
            if (active) {
                const size_t off = psz * ((qy & 3) * 4 + (qx & 3)) + (size_t)(ry0 + (qy >> 2)) * P.pstride + (rx0 + (qx >> 2));
                const uint32_t *rp = (const uint32_t *)(planes + (off & ~(size_t)3));
                const int sh = (int)(off & 3) * 8;
#pragma unroll
                for (int y = 0; y < 4; y++)
                    w[y] = LAT ? __funnelshift_r(raw[pos][2 * y], raw[pos][2 * y + 1], sh)
                               : __funnelshift_r(__ldg(rp + y * pw), __ldg(rp + y * pw + 1), sh);
            }
            if (metric == JMME_DIST_HADAMARD) {
                // 4x4 Hadamard on 16-bit lane pairs held as plain integers (hi*65536 + lo, |lane| <= 16320):
                // ordinary 32-bit add/sub act on both lanes.  The last butterfly stage pairs the two
                // lanes of one register: |lo+hi| + |lo-hi| = 2 max(|lo|,|hi|), so SATD = sum of the maxima.
                int da[4], db[4];
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    da[y] = (int)(ca[y] - __byte_perm(w[y], 0, 0x4240));   // (d0, d2)
                    db[y] = (int)(cb[y] - __byte_perm(w[y], 0, 0x4341));   // (d1, d3)
                }
                if constexpr (EXT) {
                    if (had8w) {
                        // 8x8 transform = the butterflies over the cell index (cells c^1, then c^4) followed by the
                        // 4x4 transform inside each cell; a whole cell may carry the opposite sign (abs follows)
#pragma unroll
                        for (int y = 0; y < 4; y++) {
                            int u = __shfl_xor_sync(0xFFFFFFFFu, da[y], 1), u2 = __shfl_xor_sync(0xFFFFFFFFu, db[y], 1);
                            da[y] = (cx4 & 1) ? u - da[y] : da[y] + u;
                            db[y] = (cx4 & 1) ? u2 - db[y] : db[y] + u2;
                        }
#pragma unroll
                        for (int y = 0; y < 4; y++) {
                            int u = __shfl_xor_sync(0xFFFFFFFFu, da[y], 4), u2 = __shfl_xor_sync(0xFFFFFFFFu, db[y], 4);
                            da[y] = (cy4 & 1) ? u - da[y] : da[y] + u;
                            db[y] = (cy4 & 1) ? u2 - db[y] : db[y] + u2;
                        }
                    }
                }
                int s4[4], t4[4];
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    s4[y] = da[y] + db[y];                                         // (d0+d1, d2+d3)
                    t4[y] = da[y] - db[y];                                         // (d0-d1, d2-d3)
                }
                unsigned acc = 0;
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const int *r4 = g ? t4 : s4;
                    const int u0 = r4[0] + r4[1], u1 = r4[0] - r4[1], u2 = r4[2] + r4[3], u3 = r4[2] - r4[3];
                    const int y4[4] = {u0 + u2, u0 - u2, u1 + u3, u1 - u3};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int lo = (int)(short)y4[k];
                        const int hi = (y4[k] - lo) >> 16;
                        acc += (unsigned)max(abs(lo), abs(hi));
                    }
                }
                v = active ? (int)acc : 0;           // = sum|coef| / 2 of this cell's 16 coefficients, exactly
            } else if (EXT && metric == JMME_DIST_SSE) {
                unsigned acc = 0;
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    const unsigned d = __vabsdiffu4(cw[y], w[y]);
                    acc = __dp4a(d, d, acc);
                }
                v = active ? (int)acc : 0;
            } else {
                unsigned acc = 0;
#pragma unroll
                for (int y = 0; y < 4; y++) acc = sad4(cw[y], w[y], acc);
                v = active ? (int)acc : 0;
            }
            if constexpr (EXT) {
                if (P.chroma_me) {
                    // 4:2:0: the luma vector in quarter-pel units is the chroma vector in eighth-pel units; this
                    // cell's 2x2 chroma samples of Cb and Cr, bilinear from 3x3 integer samples [STD 8.4.2.2.2]
                    const int xf = qx & 7, yf = qy & 7;
                    const int w00 = (8 - xf) * (8 - yf), w01 = xf * (8 - yf), w10 = (8 - xf) * yf, w11 = xf * yf;
                    const bool chad = metric == JMME_DIST_HADAMARD && warp < 2;   // chroma block >= 4x4: 4x4 Hadamard
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        int d[4] = {0, 0, 0, 0};
                        if (active) {
                            const uint8_t *q = P.cplanes[ref][k] + (size_t)(P.cpad + 8 * mby + 2 * cy4 + (qy >> 3)) * P.cstride +
                                               (P.cpad + 8 * mbx + 2 * cx4 + (qx >> 3));
                            int a[3][3];
#pragma unroll
                            for (int j = 0; j < 3; j++)
#pragma unroll
                                for (int i = 0; i < 3; i++) a[j][i] = __ldg(q + j * P.cstride + i);
#pragma unroll
                            for (int j = 0; j < 2; j++)
#pragma unroll
                                for (int i = 0; i < 2; i++)
                                    d[2 * j + i] = cc[k][2 * j + i] -
                                                   ((w00 * a[j][i] + w01 * a[j][i + 1] + w10 * a[j + 1][i] + w11 * a[j + 1][i + 1] + 32) >> 6);
                        }
                        if (chad) {
                            // the 4x4 chroma tile of an 8x8 luma quadrant lives in the four cells c, c^1, c^4, c^5
                            int p0 = d[1] * 65536 + d[0], p1 = d[3] * 65536 + d[2];
                            int s = p0 + p1, tt = p0 - p1;
                            int u = __shfl_xor_sync(0xFFFFFFFFu, s, 1), u2 = __shfl_xor_sync(0xFFFFFFFFu, tt, 1);
                            s = (cx4 & 1) ? u - s : s + u; tt = (cx4 & 1) ? u2 - tt : tt + u2;
                            u = __shfl_xor_sync(0xFFFFFFFFu, s, 4); u2 = __shfl_xor_sync(0xFFFFFFFFu, tt, 4);
                            s = (cy4 & 1) ? u - s : s + u; tt = (cy4 & 1) ? u2 - tt : tt + u2;
                            const int slo = (int)(short)s, shi = (s - slo) >> 16, tlo = (int)(short)tt, thi = (tt - tlo) >> 16;
                            vc += active ? max(abs(slo), abs(shi)) + max(abs(tlo), abs(thi)) : 0;   // sum|coef| / 2 of 4 coefficients
                        } else if (metric == JMME_DIST_SSE) {
                            vc += d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + d[3] * d[3];
                        } else {
                            vc += abs(d[0]) + abs(d[1]) + abs(d[2]) + abs(d[3]);
                        }
                    }
                }
                if (had8w && metric == JMME_DIST_HADAMARD) {
                    // luma: gather the 8x8 (cells c^1, c^4), then JM's (sum|coef| + 2) >> 2 per 8x8; v is sum|coef| / 2
                    v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
                    v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
                    v = P.satd_round ? (v + 1) >> 1 : v >> 1;
                    v = ((cell & 5) == 0) ? v : 0;           // one lane of the four carries the 8x8's value on
                }
                v += vc;
            }
            // sum the cells of this block: masked XOR butterfly inside the half-warp
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const int u = __shfl_xor_sync(0xFFFFFFFFu, v, o);
                if (need & o) v += u;
            }
            int cst;
            if constexpr (EXT) cst = d_wcost(P.cost_domain, P.lf[st], bitx[sx + 1] + bity[sy + 1]) + d_dscale(P.cost_domain, v);
            else cst = (int)((lf * (unsigned)(bitx[sx + 1] + bity[sy + 1])) >> 16) + v;
            if (qx == 0 && qy == 0) cst -= bonus;
            if (cst < mn) { mn = cst; best = pos; }
        }
        if constexpr (NG > 1) {                          // (cost, position) minimum over the groups
            const bool own = t <= 7 && cell == (c_blk_y[b] >> 2) * 4 + (c_blk_x[b] >> 2);
            if (active && own) s_pk[step - 1][grp][b] = ((long long)mn << 4) | best;
            __syncthreads();
            if (active) {
                long long pk = s_pk[step - 1][0][b];
#pragma unroll
                for (int g = 1; g < NG; g++) pk = min(pk, s_pk[step - 1][g][b]);
                mn = (int)(pk >> 4); best = (int)(pk & 15);
            } else {
                best = 0;
            }
        }
        mvx = ox + step * c_sp9[best][0];
        mvy = oy + step * c_sp9[best][1];
    }
    // the lane that owns the block's top-left cell publishes the result
    const bool owner = grp == 0 && t <= 7 && cell == (c_blk_y[b] >> 2) * 4 + (c_blk_x[b] >> 2);
    // every lane of the block (all in this warp) has read the packed word before the shuffles above
    if (P.int_packed && owner) P.gbest[((size_t)ref * n_mb + mb) * JMME_NBLK + b] = 0xFFFFFFFFu;
    if (active && owner) {
        BlkRes r;
        r.mvx = (int16_t)mvx; r.mvy = (int16_t)mvy; r.cost = mn;
        res[b] = r;
    }
    if (P.fused_select) {
        // one reference: nothing to choose, the record of select_ref_kernel is written here
        jmme_mbresult *o = P.out + mb;
        jmme_mbresult *q = P.out_per_ref ? P.out_per_ref + mb : nullptr;
        if (owner) {
            o->mv[b][0] = active ? (int16_t)mvx : (int16_t)0;
            o->mv[b][1] = active ? (int16_t)mvy : (int16_t)0;
            o->cost[b] = active ? mn + d_ref_cost_of(P, 0) : INT_MAX;
            o->ref_idx[b] = active ? (int8_t)0 : (int8_t)-1;
            if (q) {
                q->mv[b][0] = o->mv[b][0]; q->mv[b][1] = o->mv[b][1];
                q->cost[b] = active ? mn : INT_MAX; q->ref_idx[b] = o->ref_idx[b];
            }
        }
        if (tid < 3) {
            o->reserved[tid] = 0;
            if (q) q->reserved[tid] = 0;
        }
        if (P.n_peer_out) {                                // fused gather: this MB's record to the peers (uniform branch)
            __syncthreads();
            push_records(P, 1, [&](int) { return mb; }, tid, 128 * NG);
        }
        if (P.field_mv) {                                  // in-frame median: commit this MB (uniform branch)
            __shared__ int32_t s_cost[JMME_NBLK];
            __shared__ uint32_t s_mv[JMME_NBLK];
            __shared__ int8_t s_ref[JMME_NBLK];
            if (owner) {
                s_cost[b] = active ? mn + d_ref_cost_of(P, 0) : INT_MAX;
                s_mv[b] = active ? ((uint32_t)(uint16_t)mvx | ((uint32_t)(uint16_t)mvy << 16)) : 0u;
                s_ref[b] = active ? (int8_t)0 : (int8_t)-1;
            }
            __syncthreads();
            if (tid < 16) commit_cell(P, mb, tid, s_cost, s_mv, s_ref);
        }
    }
}

// per (MB, block): add the reference rate and keep the cheapest reference (lowest index on ties)
__global__ void select_ref_kernel(const SearchParams P)
{
    // 192 threads = 4 MBs x 48 lanes: an MB never straddles two CTAs (the commit below needs all its blocks)
    __shared__ int32_t s_cost[4][JMME_NBLK];
    __shared__ uint32_t s_mv[4][JMME_NBLK];
    __shared__ int8_t s_ref[4][JMME_NBLK];
    const int n_mb_stripe = d_n_units(P);
    const int n_mb = P.mb_w * P.mb_h;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int mbi = i / 48, b = i - mbi * 48, g = threadIdx.x / 48;
    const bool live = mbi < n_mb_stripe;
    const int mb = live ? d_unit_mb(P, mbi) : 0;
    jmme_mbresult *o = P.out + mb;
    if (live && b >= JMME_NBLK && b < JMME_NBLK + 3) {     // 3 spare lanes clear the reserved bytes
        o->reserved[b - JMME_NBLK] = 0;
        if (P.out_per_ref)
            for (int r = 0; r < P.num_refs; r++) P.out_per_ref[(size_t)r * n_mb + mb].reserved[b - JMME_NBLK] = 0;
    }
    const bool work = live && b < JMME_NBLK;
    const bool on = work && ((P.blocktype_mask >> c_blk_type[b]) & 1);
    int bc = INT_MAX, br = -1, bx = 0, by = 0;
    for (int r = 0; work && r < P.num_refs; r++) {
        BlkRes v;
        if (P.int_packed && !P.subpel) {                   // balanced integer search, no sub-pel kernel in between
            uint32_t *g = P.gbest + ((size_t)r * n_mb + mb) * JMME_NBLK + b;
            v = d_unpack_int(P, *g);
            *g = 0xFFFFFFFFu;
        } else {
            v = P.res[((size_t)r * n_mb + mb) * JMME_NBLK + b];
        }
        if (P.out_per_ref) {
            jmme_mbresult *q = P.out_per_ref + (size_t)r * n_mb + mb;
            q->mv[b][0] = on ? v.mvx : 0; q->mv[b][1] = on ? v.mvy : 0;
            q->cost[b] = on ? v.cost : INT_MAX; q->ref_idx[b] = on ? (int8_t)r : (int8_t)-1;
        }
        const int tot = v.cost + d_ref_cost_of(P, r);
        if (on && tot < bc) { bc = tot; br = r; bx = v.mvx; by = v.mvy; }
    }
    if (work) { o->mv[b][0] = (int16_t)bx; o->mv[b][1] = (int16_t)by; o->cost[b] = bc; o->ref_idx[b] = (int8_t)br; }
    if (P.n_peer_out) {                                    // fused gather: the records of this CTA's MBs to the peers
        __syncthreads();
        push_records(P, 4, [&](int k) { const int m = 4 * (int)blockIdx.x + k; return m < n_mb_stripe ? d_unit_mb(P, m) : -1; },
                     (int)threadIdx.x, 192);
    }
    if (P.field_mv) {                                      // in-frame median: commit the MBs of this CTA (uniform branch)
        if (work) {
            s_cost[g][b] = bc; s_ref[g][b] = (int8_t)br;
            s_mv[g][b] = (uint32_t)(uint16_t)bx | ((uint32_t)(uint16_t)by << 16);
        }
        __syncthreads();
        if (live && b < 16) commit_cell(P, mb, b, s_cost[g], s_mv[g], s_ref[g]);
    }
}

// stripe of the MV field -> the same offsets of up to 8 peer buffers (NVLink peer stores, 4-byte words)
struct PushArgs {
    uint32_t *dst[JMME_MAX_GPUS];
};
__global__ void __launch_bounds__(256) push_stripe_kernel(const uint32_t *__restrict__ src, PushArgs a, int n_dst,
                                                          size_t n_words)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n_words; i += (size_t)gridDim.x * 256) {
        const uint32_t v = src[i];
        for (int d = 0; d < n_dst; d++) a.dst[d][i] = v;
    }
}

}  // namespace

cudaError_t jmme_launch_push(const uint32_t *src, uint32_t *const *dst, int n_dst, size_t n_words, cudaStream_t st)
{
    PushArgs a;
    for (int d = 0; d < JMME_MAX_GPUS; d++) a.dst[d] = d < n_dst ? dst[d] : nullptr;
    const int grid = (int)std::min<size_t>((n_words + 255) / 256, 592);
    push_stripe_kernel<<<grid, 256, 0, st>>>(src, a, n_dst, n_words);
    return cudaGetLastError();
}

cudaError_t jmme_launch_subpel(const SearchParams &P, cudaStream_t st)
{
    int n_items = d_n_units(P) * P.num_refs;
    if (P.mb_list && P.pdl) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.gridDim = dim3((unsigned)n_items); cfg.blockDim = dim3(384);
        cfg.dynamicSmemBytes = 0; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
        return P.ext ? cudaLaunchKernelEx(&cfg, me_subpel_kernel<true, 3, true>, P)
                     : cudaLaunchKernelEx(&cfg, me_subpel_kernel<true, 3, false>, P);
    }
    // (the wide form on whole stripes was measured: 2.6x slower on a 9-row stripe, 2x on the frame)
    if (P.ready && !P.mb_list) {                          // early start behind the search kernel (programmatic dependent)
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.gridDim = dim3((unsigned)n_items); cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 0; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
        return P.ext ? cudaLaunchKernelEx(&cfg, me_subpel_kernel<false, 1, true>, P)
                     : cudaLaunchKernelEx(&cfg, me_subpel_kernel<false, 1, false>, P);
    }
    if (P.mb_list) {
        if (P.ext) me_subpel_kernel<true, 3, true><<<n_items, 384, 0, st>>>(P);
        else me_subpel_kernel<true, 3, false><<<n_items, 384, 0, st>>>(P);
    } else {
        if (P.ext) me_subpel_kernel<false, 1, true><<<n_items, 128, 0, st>>>(P);
        else me_subpel_kernel<false, 1, false><<<n_items, 128, 0, st>>>(P);
    }
    return cudaGetLastError();
}

cudaError_t jmme_launch_select(const SearchParams &P, cudaStream_t st)
{
    int n = d_n_units(P) * 48;
    select_ref_kernel<<<(n + 191) / 192, 192, 0, st>>>(P);
    return cudaGetLastError();
}
