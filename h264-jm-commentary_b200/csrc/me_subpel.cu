// me_subpel.cu — half- then quarter-pel refinement of all 41 blocks of a macroblock, and the
// final per-block choice of the reference picture.
//
// Stands in for JM's SubPelBlockMotionSearch with SATD ‖ HadamardSAD4x4
// (a11) and the REF_COST comparison of PartitionMotionSearch's caller (a4).  Conventions: DESIGN.md §2.
//
// Mapping: one CTA per (reference, macroblock).  A work unit is one 4x4 cell of one blocktype at
// one of the 9 (half-pel) / 8 (quarter-pel) candidate positions: the unit fetches its 4x4 reference
// samples from the quarter-pel plane selected by the candidate's fractional phase, forms the
// difference with the current MB (held in shared memory), applies the 4x4 Hadamard transform in
// registers and adds |coefficients|/2 to the (block, position) cost in shared memory.  Thread b < 41
// then adds the MV rate, takes the strict-< minimum in position order and publishes the winner.
#include "jmme_dev.cuh"

namespace {

__constant__ int8_t c_sp9[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};
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

__device__ __forceinline__ int satd16(const int (&d)[16], int satd_round)
{
    int t[16], s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int a = d[4 * i], b = d[4 * i + 1], c = d[4 * i + 2], e = d[4 * i + 3];
        int s0 = a + e, s1 = b + c, d0 = a - e, d1 = b - c;
        t[4 * i] = s0 + s1; t[4 * i + 1] = d0 + d1; t[4 * i + 2] = s0 - s1; t[4 * i + 3] = d0 - d1;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int a = t[i], b = t[4 + i], c = t[8 + i], e = t[12 + i];
        int s0 = a + e, s1 = b + c, d0 = a - e, d1 = b - c;
        // |s0+s1| + |s0-s1| = 2*max(|s0|,|s1|): the last butterfly stage folds into a max
        s += 2 * max(abs(s0), abs(s1)) + 2 * max(abs(d0), abs(d1));
    }
    return satd_round ? (s + 1) >> 1 : s >> 1;
}

__global__ void __launch_bounds__(256) me_subpel_kernel(const SearchParams P)
{
    __shared__ uint8_t s_cur[16][16];
    __shared__ int s_cost[JMME_NBLK][9];
    __shared__ int s_mvx[JMME_NBLK], s_mvy[JMME_NBLK], s_min[JMME_NBLK];
    __shared__ int s_px[JMME_NBLK], s_py[JMME_NBLK];

    const int tid = threadIdx.x;
    const int n_mb_stripe = (P.mb_row_end - P.mb_row_begin) * P.mb_w;
    const int n_mb = P.mb_w * P.mb_h;
    const int item = blockIdx.x;
    const int ref = item / n_mb_stripe;
    const int mbi = item - ref * n_mb_stripe;
    const int mby = P.mb_row_begin + mbi / P.mb_w, mbx = mbi % P.mb_w;
    const int mb = mby * P.mb_w + mbx;
    BlkRes *res = P.res + ((size_t)ref * n_mb + mb) * JMME_NBLK;
    const int bonus = (!P.rdopt && ref == 0) ? d_weighted_cost(P.lambda_factor, 16) : 0;
    const size_t psz = (size_t)P.pstride * P.pheight;
    const uint8_t *planes = P.planes[ref];

    if (tid < 64) {
        int row = tid >> 2, w = tid & 3;
        *(uint32_t *)&s_cur[row][4 * w] =
            *(const uint32_t *)(P.cur + (size_t)(16 * mby + row) * P.cur_stride + 16 * mbx + 4 * w);
    }
    if (tid < JMME_NBLK) {
        BlkRes r = res[tid];
        s_mvx[tid] = r.mvx; s_mvy[tid] = r.mvy;
        s_min[tid] = P.use_hadamard ? INT_MAX : r.cost;
        int npb = P.pred_policy == JMME_PRED_PER_BLOCK ? JMME_NBLK : 1;
        const int16_t *pr = P.pred ? P.pred + ((size_t)ref * n_mb + mb) * npb * 2 : nullptr;
        int pb = npb == 1 ? 0 : tid;
        s_px[tid] = pr ? pr[2 * pb] : 0;
        s_py[tid] = pr ? pr[2 * pb + 1] : 0;
    }

    for (int step = 2; step >= 1; step--) {
        const int pos0 = (step == 2 && P.use_hadamard) ? 0 : 1;
        for (int i = tid; i < JMME_NBLK * 9; i += 256) (&s_cost[0][0])[i] = 0;
        __syncthreads();
        // units: (pos, blocktype, cell)
        const int n_units = 9 * 7 * 16;
        for (int u = tid; u < n_units; u += 256) {
            const int pos = u / 112, rem = u - pos * 112;
            const int t = 1 + rem / 16, cell = rem & 15;
            if (pos < pos0 || !(P.blocktype_mask & (1 << t))) continue;
            const int cx4 = cell & 3, cy4 = cell >> 2;
            const int b = block_of_cell(t, cx4, cy4);
            const int qx = s_mvx[b] + step * c_sp9[pos][0], qy = s_mvy[b] + step * c_sp9[pos][1];
            const uint8_t *rp = planes + psz * ((qy & 3) * 4 + (qx & 3)) +
                                (size_t)(P.pad + 16 * mby + 4 * cy4 + (qy >> 2)) * P.pstride +
                                (P.pad + 16 * mbx + 4 * cx4 + (qx >> 2));
            int d[16];
#pragma unroll
            for (int y = 0; y < 4; y++)
#pragma unroll
                for (int x = 0; x < 4; x++)
                    d[4 * y + x] = (int)s_cur[4 * cy4 + y][4 * cx4 + x] - (int)__ldg(rp + (size_t)y * P.pstride + x);
            int v;
            if (P.use_hadamard) {
                v = satd16(d, P.satd_round);
            } else {
                v = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) v += abs(d[k]);
            }
            atomicAdd(&s_cost[b][pos], v);
        }
        __syncthreads();
        if (tid < JMME_NBLK) {
            const int b = tid;
            int mn = s_min[b], best = 0;
            const int ox = s_mvx[b], oy = s_mvy[b];
            for (int pos = pos0; pos < 9; pos++) {
                const int qx = ox + step * c_sp9[pos][0], qy = oy + step * c_sp9[pos][1];
                int c = d_weighted_cost(P.lambda_factor, d_se_bits(qx - s_px[b]) + d_se_bits(qy - s_py[b])) +
                        s_cost[b][pos];
                if (b == 0 && qx == 0 && qy == 0) c -= bonus;
                if (c < mn) { mn = c; best = pos; }
            }
            s_min[b] = mn;
            s_mvx[b] = ox + step * c_sp9[best][0];
            s_mvy[b] = oy + step * c_sp9[best][1];
        }
        __syncthreads();
    }
    if (tid < JMME_NBLK) {
        BlkRes r;
        r.mvx = (int16_t)s_mvx[tid]; r.mvy = (int16_t)s_mvy[tid]; r.cost = s_min[tid];
        res[tid] = r;
    }
}

// per (MB, block): add the reference rate and keep the cheapest reference (lowest index on ties)
__global__ void select_ref_kernel(const SearchParams P)
{
    const int n_mb_stripe = (P.mb_row_end - P.mb_row_begin) * P.mb_w;
    const int n_mb = P.mb_w * P.mb_h;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mb_stripe * 48) return;
    const int mbi = i / 48, b = i - mbi * 48;
    const int mb = P.mb_row_begin * P.mb_w + mbi;
    jmme_mbresult *o = P.out + mb;
    if (b >= JMME_NBLK) {                                  // 3 spare lanes clear the reserved bytes
        if (b < JMME_NBLK + 3) {
            o->reserved[b - JMME_NBLK] = 0;
            if (P.out_per_ref)
                for (int r = 0; r < P.num_refs; r++) P.out_per_ref[(size_t)r * n_mb + mb].reserved[b - JMME_NBLK] = 0;
        }
        return;
    }
    const bool on = (P.blocktype_mask >> c_blk_type[b]) & 1;
    int bc = INT_MAX, br = -1, bx = 0, by = 0;
    for (int r = 0; r < P.num_refs; r++) {
        BlkRes v = P.res[((size_t)r * n_mb + mb) * JMME_NBLK + b];
        if (P.out_per_ref) {
            jmme_mbresult *q = P.out_per_ref + (size_t)r * n_mb + mb;
            q->mv[b][0] = on ? v.mvx : 0; q->mv[b][1] = on ? v.mvy : 0;
            q->cost[b] = on ? v.cost : INT_MAX; q->ref_idx[b] = on ? (int8_t)r : (int8_t)-1;
        }
        const int tot = v.cost + d_ref_cost(P.lambda_factor, P.rdopt, r);
        if (on && tot < bc) { bc = tot; br = r; bx = v.mvx; by = v.mvy; }
    }
    o->mv[b][0] = (int16_t)bx; o->mv[b][1] = (int16_t)by; o->cost[b] = bc; o->ref_idx[b] = (int8_t)br;
}

}  // namespace

cudaError_t jmme_launch_subpel(const SearchParams &P, cudaStream_t st)
{
    int n_items = (P.mb_row_end - P.mb_row_begin) * P.mb_w * P.num_refs;
    me_subpel_kernel<<<n_items, 256, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t jmme_launch_select(const SearchParams &P, cudaStream_t st)
{
    int n = (P.mb_row_end - P.mb_row_begin) * P.mb_w * 48;
    select_ref_kernel<<<(n + 255) / 256, 256, 0, st>>>(P);
    return cudaGetLastError();
}
