// jmme_dev.cuh — device-side types and helpers shared by the sm_100a kernels.
// Conventions ("frozen spec") are in DESIGN.md §2; JM function names per SURVEY.md §8(a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "jmme.h"

#define JMME_NBLK 41
#define JMME_KEY_BITS 15                 // packed cost = (cost + bias) << 15 | key
#define JMME_KEY_MASK 0x7FFFu
#define JMME_NT 64                       // entries of the rate table T[bits]

// per-(ref, MB, block) intermediate result of the integer / sub-pel searches
struct __align__(8) BlkRes {
    int16_t mvx, mvy;                    // quarter-pel
    int32_t cost;                        // distortion + MV rate (no reference rate)
};

struct WaveTab;                          // wave.cuh

// launch-constant parameters of a search
struct SearchParams {
    const uint8_t *cur;                  // current luma, w16 x h16, stride cur_stride
    int cur_stride;
    int cur_h;                           // valid rows of `cur`; MB rows below replicate row cur_h-1
    const uint8_t *planes[JMME_MAX_REFS];// per reference: n_planes padded planes, plane 0 = integer
    int pstride, pheight;                // padded plane geometry
    int pad;
    int mb_w, mb_h, mb_row_begin, mb_row_end;
    int R, ncols;                        // search range, 2R+1
    int num_refs;
    int lambda_factor;
    int rdopt;                           // 0: pre-test + bonus
    int search_mode;                     // JMME_SEARCH_*
    int pred_policy;
    int blocktype_mask;
    int use_hadamard, satd_round, subpel;
    const int16_t *pred;                 // [ref][mb][nb][2] or null
    const uint16_t *spiral_key;          // [ncols*ncols]  (yoff*ncols+xoff) -> spiral index + 1
    const int16_t *spiral_xy;            // [ncols*ncols][2] spiral index -> (dx,dy)
    const uint32_t *kr0;                 // [ncols*ncols] zero predictors: ((rate + bias) << 15) + key of every candidate
                                         // (context constant, built once by the host: me_int_tb.cu KRTAB)
    BlkRes *res;                         // [ref][mb][41]
    jmme_mbresult *out;                  // [mb]
    jmme_mbresult *out_per_ref;          // [ref][mb] or null
    int fused_select;                    // 1: the sub-pel kernel writes `out` itself (one reference)
    const int *mb_list;                  // non-null: search these n_list MBs (frame MB indices) instead of the
    int n_list;                          //           stripe's rows: one step of the in-frame median wavefront
    int16_t *field_mv;                   // non-null (in-frame median): the kernel that writes `out` also commits the
    int8_t *field_ref;                   //           MB to the 4x4-granular field [4 mb_h][4 mb_w]([2])
    const WaveTab *wave_tab;             // non-null: the search kernel predicts its MB's 41 vectors itself from the
    int slice_rows;                      //           field (and writes them to `pred`) instead of reading `pred`
    int cmax;                            // the window centre pred/4 is limited to +-cmax samples (R, or jm_center: max_pred/4)
    int tune_group, tune_cluster;        // host-side launch knobs (jmme_tuning.group / .cluster, defaults resolved)
    int tune_lin;                        // 0 (jmme_tuning.table_rate): per-block rate always from the table
    int tune_split;                      // zero-predictor search with balanced task ranges (jmme_tuning.no_balance = 0)
    int pdl;                             // wavefront steps: launch with programmatic stream serialization
    // balanced task ranges of the zero-predictor search (me_int_tb.cu BAL)
    uint32_t *gbest;                     // [ref][mb][41] packed (cost + bias) << 15 | key minima, 0xFFFFFFFF between searches
    int int_packed;                      // 1: the integer search leaves its result in gbest (not in res); the first kernel
                                         //    that consumes it — sub-pel, else reference selection — decodes and resets
    // early start of the sub-pel kernel (me_int_tb.cu / me_subpel.cu): the search kernel raises ready[ref][mb] when the
    // integer result of an MB is in `res`, the sub-pel kernel — a programmatic dependent that starts while the search
    // kernel's last round is still running — waits for its MB's flag instead of the kernel boundary and lowers it again
    int *ready;
    int pair_row;                        // > 0 (items of 4 MBs): MB rows from here on are dealt out as items of 2 MBs
    jmme_mbresult *peer_out[JMME_MAX_GPUS];  // fused gather: the kernel that writes a record of `out` also stores it into
    int n_peer_out;                      //               the same offset of these (peer-mapped) buffers
    jmme_mbresult *mc_out;               // non-null (then n_peer_out = -1: "fused gather on", no peer list): one multimem.st per
                                         //               word through this NVLS multicast mapping reaches every rank's field
    // ---- ABI 4: cost domain, per-stage metrics, 8x8 Hadamard, chroma ME ----
    int cost_domain;                     // 0: D + (lf*bits >> 16)   1: (D << 5) + lf*bits
    int metric[3], lf[3];                // JMME_DIST_* and lambda factor of the integer / half-pel / quarter-pel stage
    int ext;                             // 1: anything beyond the legacy path is on (domain 1, SSE, 8x8, chroma, or metrics
                                         //    that restart the quarter-pel stage): kernels take their general form
    int t8, chroma_me;
    const uint8_t *cplanes[JMME_MAX_REFS][2];   // padded integer chroma planes (Cb, Cr) of every reference
    int cstride, cpad;
    const uint8_t *cur_c[2];             // current chroma, w16/2 x h16/2
    int cur_cs;
};

// copy the records of n_rec MBs (mb_of(i) = frame MB index of record i) from P.out to every peer buffer; the
// calling threads wrote those records themselves and have passed a barrier since
template <class F>
__device__ __forceinline__ void push_records(const SearchParams &P, int n_rec, F mb_of, int tid, int nthreads)
{
    constexpr int RW = sizeof(jmme_mbresult) / 4;
    for (int j = tid; j < n_rec * RW; j += nthreads) {     // one word per thread: read once, store to every peer
        const int rec = j / RW, w = j - rec * RW, mb = mb_of(rec);
        if (mb < 0) continue;
        const uint32_t v = __ldcg((const uint32_t *)(P.out + mb) + w);
        if (P.mc_out)                                      // the switch replicates the store to every rank
            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"((uint32_t *)(P.mc_out + mb) + w), "r"(v) : "memory");
        for (int p = 0; p < P.n_peer_out; p++) ((uint32_t *)(P.peer_out[p] + mb))[w] = v;
    }
}

// ---- host side: per-(kernel instantiation, device) launch state -------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize is process-wide per (kernel, device): it is raised once to the
// device's opt-in maximum (never lowered, so host threads with different search ranges cannot undo each
// other), and the occupancy of the last (bytes) is cached per device.  One static KernelState per template
// instantiation of a launch function.
#include <mutex>
#define JMME_MAX_DEVICES 64
struct KernelState {
    std::mutex mu;
    bool attr_set[JMME_MAX_DEVICES] = {};
    size_t bytes[JMME_MAX_DEVICES] = {};
    int occ[JMME_MAX_DEVICES] = {};
};
template <class Kern>
cudaError_t jmme_kernel_occupancy(Kern kern, KernelState &ks, int threads, size_t bytes, int *occ_out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= JMME_MAX_DEVICES) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(ks.mu);
    if (!ks.attr_set[dev]) {
        int optin = 0;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e != cudaSuccess) return e;
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
        if (e != cudaSuccess) return e;
        ks.attr_set[dev] = true;
    }
    if (ks.bytes[dev] != bytes || ks.occ[dev] == 0) {
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, bytes);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        ks.bytes[dev] = bytes; ks.occ[dev] = occ;
    }
    *occ_out = ks.occ[dev];
    return cudaSuccess;
}
// name of the integer-search kernel instantiation a launch function picked (jmme_last_kernel); per host thread
char *jmme_kernel_name_buf();
#define JMME_KNAME_LEN 320

// Loops in front of a CTA barrier must have a warp-uniform trip count (`for (x0 = 0; x0 < n; x0 += 32) { x = x0 + lane;
// if (x >= n) continue; ... }`, not `for (x = lane; x < n; x += 32)`).  ptxas 12.9 was seen to drop the reconvergence
// point (BSSY/BSYNC) of a lane-strided loop — the window expansion of me_int_tb.cu, once an epilogue behind the item
// loop turned its `break` into a branch: the lanes without a second trip ran ahead, reached BAR.SYNC while the warp
// was still diverged (bar.sync is warp-aligned: the warp then counts as arrived) and the barrier released before the
// other lanes had written their window columns.  An explicit __syncwarp() in front of the barrier does not help:
// ptxas turns it into a NOP where its analysis says "converged".  Found by the whole-frame parity tests (every item
// after a CTA's first was wrong), verified in the SASS.
// Programmatic dependent launch (sm_90+).  pdl_trigger: the next kernel of the stream may start its prologue;
// pdl_wait: results of the previous kernel are complete and visible from here on.  Both are no-ops for a
// kernel launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// the MBs one launch works on: a stripe of MB rows, or an explicit list
__host__ __device__ inline int d_n_units(const SearchParams &P)
{
    return P.mb_list ? P.n_list : (P.mb_row_end - P.mb_row_begin) * P.mb_w;
}
__device__ __forceinline__ int d_unit_mb(const SearchParams &P, int i)
{
    return P.mb_list ? P.mb_list[i] : P.mb_row_begin * P.mb_w + i;
}

__device__ __forceinline__ int d_se_bits(int v)
{
    // signed Exp-Golomb length: 1 for 0, else 2*floor(log2|v|)+3
    int a = abs(v);
    return a ? 2 * (31 - __clz(a)) + 3 : 1;
}
__device__ __forceinline__ int d_ue_bits(int r) { return 2 * (31 - __clz(r + 1)) + 1; }
__device__ __forceinline__ int d_weighted_cost(int f, int bits) { return (int)(((long long)f * bits) >> 16); }
__device__ __forceinline__ int d_ref_cost(int f, int rdopt, int ref)
{
    if (rdopt) return d_weighted_cost(f, d_ue_bits(ref));
    return ref ? (int)((2ll * f) >> 16) : 0;
}
// the same in either cost domain: domain 1 = JM >= 12 scaled-up costs, rate not truncated
#define JMME_SCALEUP_BITS 5
__device__ __forceinline__ int d_wcost(int domain, int f, int bits) { return domain ? f * bits : d_weighted_cost(f, bits); }
__device__ __forceinline__ int d_dscale(int domain, int d) { return domain ? d << JMME_SCALEUP_BITS : d; }
// reference rate with the lambda factor of the last stage that ran
__device__ __forceinline__ int d_ref_cost_of(const SearchParams &P, int ref)
{
    const int f = P.lf[P.subpel ? 2 : 0];
    if (P.rdopt) return d_wcost(P.cost_domain, f, d_ue_bits(ref));
    return ref ? (P.cost_domain ? 2 * f : (int)((2ll * f) >> 16)) : 0;
}
__device__ __forceinline__ int d_clamp(int v, int lo, int hi) { return min(max(v, lo), hi); }
// Predictor components as the kernels use them: the host entry points refuse |pred| > JMME_MAX_PRED_QPEL, the
// device-pointer entry points cannot look, so every load clamps (keeps se_bits sums inside the rate tables).
__device__ __forceinline__ int d_pred(int v) { return d_clamp(v, -JMME_MAX_PRED_QPEL, JMME_MAX_PRED_QPEL); }

// spiral index k (0 = centre) -> (dx, dy), the inverse of the host's spiral_index(): ring l = max(|dx|, |dy|) starts
// at (2l-1)^2 with its top/bottom rows interleaved (x = -l+1 .. l-1), then its left/right columns (y = -l .. l).
// Arithmetic instead of the spiral_xy table: the result write of a search kernel has no dependent global load.
__device__ __forceinline__ void d_spiral_xy(int k, int &dx, int &dy)
{
    if (k == 0) { dx = dy = 0; return; }
    int s = (int)sqrtf((float)k);
    s -= (s * s > k);
    s += ((s + 1) * (s + 1) <= k);
    const int l = (s + 1) >> 1, w = 2 * l - 1, off = k - w * w;
    if (off < 2 * w) { dx = (off >> 1) - l + 1; dy = (off & 1) ? l : -l; }
    else { const int o2 = off - 2 * w; dy = (o2 >> 1) - l; dx = (o2 & 1) ? l : -l; }
}

// integer result of (ref, mb, block) left by a balanced search: decode the packed minimum (zero predictors: window
// centre (0,0)) and reset the word for the next search.  One caller per word, or callers of one warp with the reset
// issued by one of them after all have read (me_subpel.cu).
__device__ __forceinline__ BlkRes d_unpack_int(const SearchParams &P, unsigned v)
{
    BlkRes r;
    int dx = 0, dy = 0;
    const unsigned key = v & JMME_KEY_MASK;
    if (key) d_spiral_xy((int)key - 1, dx, dy);
    r.mvx = (int16_t)(4 * dx); r.mvy = (int16_t)(4 * dy);
    r.cost = (int)(v >> JMME_KEY_BITS) - (P.rdopt ? 0 : d_weighted_cost(P.lambda_factor, 16));
    return r;
}

// 4 absolute byte differences summed and accumulated: one VABSDIFF4.U8.ACC
__device__ __forceinline__ unsigned sad4(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- ME-only mode decision, shared by commit_kernel and the in-frame median epilogues ----
// index base of blocktype t in the result order: 0,0,1,3,5,9,17,25
__device__ __forceinline__ int blk_base_of(int t) { return (int)((0x1911090503010000ull >> (8 * t)) & 0xFF); }

// cost[41] of one MB -> low 2 bits = 0..2 for 16x16 / 16x8 / 8x16, 3 = P8x8; bits 4+4q.. = sub-type of 8x8
// number q.  (One packed word, no local arrays: this runs on the serial path of the wavefront.)
__device__ inline int mb_mode(const int32_t *cost, int mask)
{
    const long long INF = 0x7FFFFFFFFFFFFFFFll;
    const long long J0 = (mask >> 1) & 1 ? (long long)cost[0] : INF;
    const long long J1 = (mask >> 2) & 1 ? (long long)cost[1] + cost[2] : INF;
    const long long J2 = (mask >> 3) & 1 ? (long long)cost[3] + cost[4] : INF;
    long long J3 = 0;
    int subs = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        long long bq = INF;
        int st = 0;
#pragma unroll
        for (int t = 4; t <= 7; t++) {
            if (!((mask >> t) & 1)) continue;
            const int bw = t <= 5 ? 8 : 4, bh = (t == 4 || t == 6) ? 8 : 4, nbx = 16 / bw;
            long long s = 0;
#pragma unroll
            for (int sy = (q >> 1) * (8 / bh); sy < ((q >> 1) + 1) * (8 / bh); sy++)      // blocks of type t in 8x8 q
#pragma unroll
                for (int sx = (q & 1) * (8 / bw); sx < ((q & 1) + 1) * (8 / bw); sx++) s += cost[blk_base_of(t) + sy * nbx + sx];
            if (s < bq) { bq = s; st = t; }
        }
        subs |= st << (4 + 4 * q);
        if (J3 != INF) J3 = bq == INF ? INF : J3 + bq;
    }
    int md = 0;                                            // lower mode wins ties
    long long best = J0;
    if (J1 < best) { best = J1; md = 1; }
    if (J2 < best) { best = J2; md = 2; }
    if (J3 < best) md = 3;
    return md | subs;
}
// field cell (cx4, cy4) of an MB decided as `mode`: the block that covers it
__device__ inline int cell_block(int mode, int cx4, int cy4)
{
    const int md = mode & 3;
    const int t = md == 3 ? (mode >> (4 + 4 * (2 * (cy4 >> 1) + (cx4 >> 1)))) & 15 : md + 1;
    const int lw = (t == 1 || t == 2) ? 2 : (t <= 5 ? 1 : 0), lh = (t == 1 || t == 3) ? 2 : ((t == 2 || t == 4 || t == 6) ? 1 : 0);
    return blk_base_of(t) + ((cy4 >> lh) << (2 - lw)) + (cx4 >> lw);      // (4cy4 / bh) * (16 / bw) + 4cx4 / bw
}
// commit field cell `cell` (0..15) of MB `mb` from the MB's 41 (cost, packed mv, ref) triples
__device__ inline void commit_cell(const SearchParams &P, int mb, int cell, const int32_t *cost, const uint32_t *mv,
                                   const int8_t *ref)
{
    const int cx4 = cell & 3, cy4 = cell >> 2;
    const int blk = cell_block(mb_mode(cost, P.blocktype_mask), cx4, cy4);
    const int mby = mb / P.mb_w, mbx = mb - mby * P.mb_w;
    const size_t o = (size_t)(4 * mby + cy4) * (4 * P.mb_w) + 4 * mbx + cx4;
    *(uint32_t *)(P.field_mv + 2 * o) = mv[blk];
    P.field_ref[o] = ref[blk];
}

// block geometry, result order (blocktype 1..7, raster inside the MB)
__device__ __constant__ const uint8_t c_blk_x[JMME_NBLK] = {
    0, 0, 0, 0, 8, 0, 8, 0, 8, 0, 8, 0, 8, 0, 8, 0, 8,
    0, 4, 8, 12, 0, 4, 8, 12, 0, 4, 8, 12, 0, 4, 8, 12, 0, 4, 8, 12, 0, 4, 8, 12};
__device__ __constant__ const uint8_t c_blk_y[JMME_NBLK] = {
    0, 0, 8, 0, 0, 0, 0, 8, 8, 0, 0, 4, 4, 8, 8, 12, 12,
    0, 0, 0, 0, 8, 8, 8, 8, 0, 0, 0, 0, 4, 4, 4, 4, 8, 8, 8, 8, 12, 12, 12, 12};
__device__ __constant__ const uint8_t c_blk_w[JMME_NBLK] = {
    16, 16, 16, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};
__device__ __constant__ const uint8_t c_blk_h[JMME_NBLK] = {
    16, 8, 8, 16, 16, 8, 8, 8, 8, 4, 4, 4, 4, 4, 4, 4, 4,
    8, 8, 8, 8, 8, 8, 8, 8, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};
__device__ __constant__ const uint8_t c_blk_type[JMME_NBLK] = {
    1, 2, 2, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5,
    6, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7};
