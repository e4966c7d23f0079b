// jmme_api.cu — the C ABI of include/jmme.h on top of the sm_100a kernels.
//
// Host side of the drop-in boundary: context life cycle, device memory, stream ordering, the
// stripe (MB-row) partition, the single-process multi-GPU mode (one sub-context per device, MV
// field gathered to the first device over NVLink peer copies) and parameter checking that
// mirrors the oracle's error behaviour.  There is NO CPU fallback: every compute entry point
// fails with JMME_ERR_NODEVICE / JMME_ERR_CUDA when no B200 is usable.
#include <cuda_runtime.h>

#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "jmme_dev.cuh"
#include "wave.cuh"

const WaveTab &jmme_wave_tab();
char *jmme_kernel_name_buf()
{
    static thread_local char buf[JMME_KNAME_LEN];
    return buf;
}

cudaError_t jmme_launch_me_int(const SearchParams &P, int num_sms, int variant, cudaStream_t st);
cudaError_t jmme_launch_me_full(const SearchParams &P, cudaStream_t st);
bool jmme_me_int_balanced(const SearchParams &P, int variant, int num_sms, bool forced);      // me_int_tb.cu
bool jmme_me_int_raises_flags(const SearchParams &P, int variant);                            // me_int_tb.cu
cudaError_t jmme_launch_interp(const uint8_t *src, int w_in, int h_in, int stride, int pad, int ps, int ph,
                               int n_planes, uint8_t *out, int y_begin, int y_end, cudaStream_t st);
cudaError_t jmme_launch_pad_cur(const uint8_t *src, int w_in, int h_in, int stride, int w16, int h16, uint8_t *dst,
                                cudaStream_t st);
cudaError_t jmme_launch_subpel(const SearchParams &P, cudaStream_t st);
cudaError_t jmme_launch_select(const SearchParams &P, cudaStream_t st);
cudaError_t jmme_launch_predict(const int16_t *mv4, const int8_t *ref4, int mb_w, int mb_h, int num_refs, int16_t *pred,
                                cudaStream_t st);
cudaError_t jmme_launch_commit(const jmme_mbresult *res, int mb_w, int mb_h, int mask, int16_t *mv4, int8_t *ref4,
                               uint8_t *mode, cudaStream_t st);
cudaError_t jmme_launch_wave_step(const int *cur, int n_cur, int mb_w, int mb_h, int num_refs, int slice_rows,
                                  const int16_t *mv4, const int8_t *ref4, int16_t *pred, cudaStream_t st);
cudaError_t jmme_launch_pad_plane(const uint8_t *src, int w, int h, int stride, int pad, int pw, int ph, uint8_t *dst,
                                  cudaStream_t st);
struct BipredArgs {                      // me_bipred.cu
    const uint8_t *planes_l1;
    const jmme_mbresult *l0, *l1;
    const int16_t *pred0, *pred1;
    const int16_t *spiral_xy;
    int range, iterations, npb, n_planes;
    jmme_bipred *out;
    int *err;
};
cudaError_t jmme_launch_bipred(const SearchParams &P, const BipredArgs &A, cudaStream_t st);
cudaError_t jmme_launch_push(const uint32_t *src, uint32_t *const *dst, int n_dst, size_t n_words, cudaStream_t st);

struct jmme_ctx {
    jmme_params p;
    int w16, h16, mb_w, mb_h, pad, pstride, pheight, lambda_factor, n_planes, ncols, ncand;
    int max_pred, cmax;                   // |pred| limit of the host entry points; limit of the window centre (samples)
    int device, num_sms;
    int metric[3], lf[3], ext;            // per-stage JMME_DIST_* / lambda factor in the context's cost domain; ext: general kernels
    int cpad, cstride, cheight;           // chroma ME: padded integer chroma planes (w16/2 + 2 cpad) x (h16/2 + 2 cpad)
    uint8_t *d_cplanes[JMME_MAX_REFS][2];
    bool cref_set[JMME_MAX_REFS], cur_c_set;
    uint8_t *d_cur_c[2];                  // current chroma, w16/2 x h16/2
    uint8_t *d_craw;                      // staging for host chroma uploads (two components)
    // bi-predictive refinement (allocated by the first jmme_set_reference_l1)
    uint8_t *d_planes_l1, *d_raw_l1;
    bool l1_set;
    jmme_mbresult *d_bi_l0, *d_bi_l1;
    int16_t *d_bi_pred0, *d_bi_pred1, *d_bi_spiral;
    jmme_bipred *d_bi_out;
    int *d_bi_err;
    jmme_tuning tune;                     // launch knobs with the defaults resolved (jmme_set_tuning)
    char last_kernel[JMME_KNAME_LEN];     // integer-search kernel instantiation of the last search
    cudaStream_t stream;
    cudaStream_t copy_stream;             // host->device copy of the current picture, overlaps the plane kernel
    cudaEvent_t ev_copy;
    cudaStream_t part_stream[4];          // the pipelined host path searches the stripe in up to 4 parts
    cudaEvent_t ev_ref;
    // async_reference + pipelined host path: jmme_set_reference queues only the first chunk of the picture (the rows
    // the first part of the search needs); the other chunks are uploaded and interpolated by jmme_search_frame,
    // interleaved with the parts of the current picture, so the first search kernel waits for a fraction of both
    // pictures instead of the whole reference
    const uint8_t *pend_luma[JMME_MAX_REFS];
    int pend_stride[JMME_MAX_REFS];
    int pend_next[JMME_MAX_REFS];         // first chunk not queued yet; 0 = nothing pending
    cudaEvent_t ev_chunk[4];
    cudaEvent_t ev_search;                // end of the last in-frame median search (orders jmme_get_predictors)
    uint8_t *d_raw;                       // staging for the raw current picture (width x height)
    uint8_t *d_raw_ref[JMME_MAX_REFS];    // staging for the raw reference pictures
    uint8_t *d_planes[JMME_MAX_REFS];
    bool ref_set[JMME_MAX_REFS];
    uint8_t *d_cur16;
    int16_t *d_pred;
    uint16_t *d_spiral_key;
    int16_t *d_spiral_xy;
    uint32_t *d_kr0;                      // zero-predictor rate + key table of the context (SearchParams::kr0)
    int *d_ready;                         // per-(ref, MB) flags of the early sub-pel start (SearchParams::ready), all 0
    uint32_t *d_gbest;                    // packed minima of a balanced integer search (SearchParams::gbest), all 0xFFFFFFFF
    bool ready_dirty;                     // the same for d_ready
    bool gbest_dirty;                     // a failed search may have left words behind: cleared before the next one
    BlkRes *d_res;
    jmme_mbresult *d_out, *d_out_per_ref;
    long long launches;
    char err[256];
    int n_sub;                            // >0: this is a multi-GPU parent, work lives in sub[]
    jmme_ctx *sub[JMME_MAX_GPUS];
    cudaEvent_t ev_done;
    int profiling;                        // bracket kernels with events
    cudaEvent_t ev_prof[4][2];            // [interp, me_int, me_subpel, select][begin, end]
    bool prof_valid[4];
    // JMME_PRED_MEDIAN: the field committed so far, and the MBs of every wavefront step of the stripe
    int16_t *d_fmv;
    int8_t *d_fref;
    WaveTab *d_wave_tab;                  // neighbour-source table of the predictor code (wave.cuh)
    int *d_wave;                          // MB indices, step after step
    int *wave_off;                        // [n_steps + 1] offsets into d_wave (host)
    int n_steps;
    bool searched;                        // d_pred holds the predictors of a finished median search
    void *peer_fields[JMME_MAX_GPUS];     // jmme_set_peer_fields_dev
    int n_peer_fields;
    void *mc_field;                       // jmme_set_multicast_field_dev
    bool dev_call;                        // inside jmme_search_frame_dev (the fused gather applies to that call only)
};

namespace {

int fail(jmme_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (c) {
        if (e != cudaSuccess)
            snprintf(c->err, sizeof c->err, "%s: %s", what, cudaGetErrorString(e));
        else
            snprintf(c->err, sizeof c->err, "%s", what);
    }
    return code;
}
#define CU(c, call)                                                        \
    do {                                                                   \
        cudaError_t e_ = (call);                                           \
        if (e_ != cudaSuccess) return fail(c, JMME_ERR_CUDA, #call, e_);   \
    } while (0)

// replication border: the window centre moves up to cmax samples from the MB, the window R more, 16 for the MB and
// the sub-pel / 6-tap margin (cmax = R unless jm_center lets the centre follow the predictor)
int pad_for(int R, int cmax) { return (cmax + R + 16 + 15) & ~15; }

// jmme_tuning with every 0 replaced by the measured default
void resolve_tuning(const jmme_tuning *in, jmme_tuning *out)
{
    jmme_tuning t;
    memset(&t, 0, sizeof t);
    if (in) t = *in;
    if (t.group < 0) t.group = 0;                 // MBs per work item of the zero-predictor kernel (0: 4 balanced, else 2)
    if (t.cluster <= 0) t.cluster = 4;            // largest cluster of a wavefront step (1 = none)
    t.pipe_parts = t.pipe_parts <= 0 ? 3 : std::min(t.pipe_parts, 4);
    *out = t;
}

// rows of `width` bytes host -> device; one linear copy when the source rows are contiguous
cudaError_t upload_rows(uint8_t *dst, const uint8_t *src, int stride, int width, int rows, cudaStream_t st)
{
    if (stride == width) return cudaMemcpyAsync(dst, src, (size_t)width * rows, cudaMemcpyHostToDevice, st);
    return cudaMemcpy2DAsync(dst, width, src, stride, width, rows, cudaMemcpyHostToDevice, st);
}

// JM spiral order generated from its closed form: ring l = max(|dx|,|dy|), then the
// top/bottom rows interleaved, then the left/right columns interleaved.
int spiral_index(int dx, int dy)
{
    if (!dx && !dy) return 0;
    int l = std::max(std::abs(dx), std::abs(dy)), base = (2 * l - 1) * (2 * l - 1);
    if (std::abs(dy) == l && std::abs(dx) < l) return base + 2 * (dx + l - 1) + (dy > 0);
    return base + 2 * (2 * l - 1) + 2 * (dy + l) + (dx > 0);
}

const int kQp2Quant[40] = {1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 23,
                           25, 29, 32, 36, 40, 45, 51, 57, 64, 72, 81, 91};

void free_device(jmme_ctx *c)
{
    if (c->device >= 0) cudaSetDevice(c->device);
    cudaFree(c->d_raw);
    for (int r = 0; r < JMME_MAX_REFS; r++) {
        cudaFree(c->d_planes[r]); cudaFree(c->d_raw_ref[r]); cudaFree(c->d_cplanes[r][0]); cudaFree(c->d_cplanes[r][1]);
    }
    cudaFree(c->d_cur_c[0]); cudaFree(c->d_cur_c[1]); cudaFree(c->d_craw);
    cudaFree(c->d_planes_l1); cudaFree(c->d_raw_l1); cudaFree(c->d_bi_l0); cudaFree(c->d_bi_l1); cudaFree(c->d_bi_pred0);
    cudaFree(c->d_bi_pred1); cudaFree(c->d_bi_spiral); cudaFree(c->d_bi_out); cudaFree(c->d_bi_err);
    cudaFree(c->d_cur16); cudaFree(c->d_pred); cudaFree(c->d_spiral_key); cudaFree(c->d_spiral_xy); cudaFree(c->d_kr0); cudaFree(c->d_gbest); cudaFree(c->d_ready);
    cudaFree(c->d_res); cudaFree(c->d_out); cudaFree(c->d_out_per_ref);
    cudaFree(c->d_fmv); cudaFree(c->d_fref); cudaFree(c->d_wave); cudaFree(c->d_wave_tab);
    free(c->wave_off);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 2; j++)
            if (c->ev_prof[i][j]) cudaEventDestroy(c->ev_prof[i][j]);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->ev_ref) cudaEventDestroy(c->ev_ref);
    for (int i = 0; i < 4; i++)
        if (c->ev_chunk[i]) cudaEventDestroy(c->ev_chunk[i]);
    if (c->ev_search) cudaEventDestroy(c->ev_search);
    for (int i = 0; i < 4; i++)
        if (c->part_stream[i]) cudaStreamDestroy(c->part_stream[i]);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
}

int validate(const jmme_params *p)
{
    if (p->width <= 0 || p->height <= 0 || p->search_range < 1 || p->search_range > JMME_MAX_SEARCH_RANGE ||
        p->num_refs < 1 || p->num_refs > JMME_MAX_REFS || (p->blocktype_mask & ~JMME_MASK_ALL) ||
        !(p->blocktype_mask & JMME_MASK_ALL) || p->qp < 0 || p->qp > 51 || p->lambda_factor < 0 ||
        p->search_mode < 0 || p->search_mode > 1 || p->pred_policy < 0 || p->pred_policy > 3 || p->satd_round < 0 ||
        p->satd_round > 1 || p->n_gpus < 0 || p->n_gpus > JMME_MAX_GPUS || p->slice_rows < 0 || p->cost_domain < 0 ||
        p->cost_domain > 1 || p->me_distortion < 0 || p->me_distortion > 1 || p->transform8x8 < 0 || p->transform8x8 > 1 ||
        p->chroma_me < 0 || p->chroma_me > 1 || p->jm_center < 0 || p->jm_center > 1 ||
        (p->max_pred_qpel && (p->max_pred_qpel < 4 || p->max_pred_qpel > JMME_MAX_PRED_QPEL)))
        return JMME_ERR_PARAM;
    if (p->me_distortion &&
        (p->me_distortion_fpel < 0 || p->me_distortion_fpel > 2 || p->me_distortion_hpel < 0 || p->me_distortion_hpel > 2 ||
         p->me_distortion_qpel < 0 || p->me_distortion_qpel > 2))
        return JMME_ERR_PARAM;
    // the integer stage builds its surfaces from per-pixel sums (SAD or SSE): no Hadamard there (as in the oracle)
    if (p->me_distortion && p->me_distortion_fpel == JMME_DIST_HADAMARD) return JMME_ERR_UNSUPPORTED;
    if (p->chroma_me && !p->subpel) return JMME_ERR_PARAM;          // chroma enters at the sub-pel stages
    return JMME_OK;
}

int create_single(jmme_ctx **out, const jmme_params *p, int device)
{
    jmme_ctx *c = new (std::nothrow) jmme_ctx();
    if (!c) return JMME_ERR_NOMEM;
    memset(c, 0, sizeof *c);
    c->p = *p;
    c->device = device;
    c->w16 = (p->width + 15) & ~15; c->h16 = (p->height + 15) & ~15;
    c->mb_w = c->w16 / 16; c->mb_h = c->h16 / 16;
    if (c->p.mb_row_end == 0) c->p.mb_row_end = c->mb_h;
    if (c->p.mb_row_begin < 0 || c->p.mb_row_end > c->mb_h || c->p.mb_row_begin >= c->p.mb_row_end) {
        delete c; return JMME_ERR_PARAM;
    }
    if (p->pred_policy == JMME_PRED_MEDIAN) {     // stripes start and end on slice boundaries
        const int k = p->slice_rows ? p->slice_rows : c->mb_h;
        if (c->p.mb_row_begin % k || (c->p.mb_row_end % k && c->p.mb_row_end != c->mb_h)) { delete c; return JMME_ERR_PARAM; }
    }
    c->max_pred = p->max_pred_qpel ? p->max_pred_qpel : JMME_MAX_PRED_QPEL;
    // centre limit: +-R, or (JM: rdopt = 1 does not clamp) as far as the largest accepted predictor reaches
    c->cmax = std::max((p->jm_center && p->rdopt) ? c->max_pred / 4 : p->search_range, p->search_range);
    c->pad = pad_for(p->search_range, c->cmax);
    c->pstride = c->w16 + 2 * c->pad; c->pheight = c->h16 + 2 * c->pad;
    {
        // per-stage metric and lambda factor: an SSE stage works with lambda^2; cost domain 1 scales
        // lambda by 32 instead of 65536 and leaves the rate untruncated
        const int lf16 = p->lambda_factor ? p->lambda_factor : jmme_lambda_factor(p->qp, p->rdopt);
        if (lf16 > (96 << 16)) { delete c; return JMME_ERR_PARAM; }
        const int q = std::min(std::max(p->qp - 12, 0), 39);
        const double lambda = p->lambda_factor ? (double)p->lambda_factor / 65536.0
                                               : (p->rdopt ? std::sqrt(0.85 * std::pow(2.0, q / 3.0)) : (double)kQp2Quant[q]);
        c->metric[0] = p->me_distortion ? p->me_distortion_fpel : JMME_DIST_SAD;
        c->metric[1] = p->me_distortion ? p->me_distortion_hpel : (p->use_hadamard ? JMME_DIST_HADAMARD : JMME_DIST_SAD);
        c->metric[2] = p->me_distortion ? p->me_distortion_qpel : (p->use_hadamard ? JMME_DIST_HADAMARD : JMME_DIST_SAD);
        for (int st = 0; st < 3; st++) {
            const double l = c->metric[st] == JMME_DIST_SSE ? lambda * lambda : lambda;
            c->lf[st] = (int)((p->cost_domain ? 32.0 : 65536.0) * l + 0.5);
        }
        c->lambda_factor = c->lf[0];
        // the legacy kernels cover: domain 0, SAD integer stage, sub-pel stages both SAD or both Hadamard 4x4, luma only
        c->ext = p->cost_domain || p->transform8x8 || p->chroma_me || c->metric[0] != JMME_DIST_SAD ||
                 c->metric[1] == JMME_DIST_SSE || c->metric[2] == JMME_DIST_SSE || c->metric[1] != c->metric[2];
    }
    c->cpad = c->pad / 2; c->cstride = c->w16 / 2 + 2 * c->cpad; c->cheight = c->h16 / 2 + 2 * c->cpad;
    c->n_planes = p->subpel ? 16 : 1;
    c->ncols = 2 * p->search_range + 1; c->ncand = c->ncols * c->ncols;
    resolve_tuning(nullptr, &c->tune);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device >= ndev) { delete c; return JMME_ERR_NODEVICE; }
    int rc = JMME_OK;
#define CUC(call)                                                                      \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) { rc = fail(c, JMME_ERR_CUDA, #call, e_); goto bad; }   \
    } while (0)
    {
        const size_t n_mb = (size_t)c->mb_w * c->mb_h;
        const size_t psz = (size_t)c->pstride * c->pheight;
        std::vector<uint16_t> key(c->ncand);
        std::vector<int16_t> xy(2 * (size_t)c->ncand);
        const int R = p->search_range;
        CUC(cudaSetDevice(device));
        CUC(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
        CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CUC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CUC(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
        for (int i = 0; i < 4; i++) CUC(cudaStreamCreateWithFlags(&c->part_stream[i], cudaStreamNonBlocking));
        CUC(cudaEventCreateWithFlags(&c->ev_ref, cudaEventDisableTiming));
        for (int i = 0; i < 4; i++) CUC(cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&c->ev_search, cudaEventDisableTiming));
        CUC(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
        CUC(cudaMalloc(&c->d_raw, (size_t)p->width * p->height));
        for (int r = 0; r < p->num_refs; r++) {
            CUC(cudaMalloc(&c->d_planes[r], psz * c->n_planes));
            CUC(cudaMalloc(&c->d_raw_ref[r], (size_t)p->width * p->height));
        }
        CUC(cudaMalloc(&c->d_cur16, (size_t)c->w16 * c->h16));
        if (p->chroma_me) {
            for (int r = 0; r < p->num_refs; r++)
                for (int k = 0; k < 2; k++) CUC(cudaMalloc(&c->d_cplanes[r][k], (size_t)c->cstride * c->cheight));
            for (int k = 0; k < 2; k++) CUC(cudaMalloc(&c->d_cur_c[k], (size_t)(c->w16 / 2) * (c->h16 / 2)));
            CUC(cudaMalloc(&c->d_craw, 2 * (size_t)((p->width + 1) / 2) * ((p->height + 1) / 2)));
        }
        CUC(cudaMalloc(&c->d_pred, sizeof(int16_t) * 2 * JMME_NBLK * n_mb * p->num_refs));
        CUC(cudaMalloc(&c->d_res, sizeof(BlkRes) * JMME_NBLK * n_mb * p->num_refs));
        CUC(cudaMemset(c->d_res, 0, sizeof(BlkRes) * JMME_NBLK * n_mb * p->num_refs));
        CUC(cudaMalloc(&c->d_out, sizeof(jmme_mbresult) * n_mb));
        CUC(cudaMemset(c->d_out, 0, sizeof(jmme_mbresult) * n_mb));
        CUC(cudaMalloc(&c->d_out_per_ref, sizeof(jmme_mbresult) * n_mb * p->num_refs));
        CUC(cudaMalloc(&c->d_spiral_key, sizeof(uint16_t) * c->ncand));
        CUC(cudaMalloc(&c->d_spiral_xy, sizeof(int16_t) * 2 * c->ncand));
        for (int dy = -R; dy <= R; dy++)
            for (int dx = -R; dx <= R; dx++) {
                int k = spiral_index(dx, dy);
                key[(size_t)(dy + R) * c->ncols + (dx + R)] = (uint16_t)(k + 1);
                xy[2 * (size_t)k] = (int16_t)dx; xy[2 * (size_t)k + 1] = (int16_t)dy;
            }
        CUC(cudaMemcpy(c->d_spiral_key, key.data(), sizeof(uint16_t) * c->ncand, cudaMemcpyHostToDevice));
        CUC(cudaMemcpy(c->d_spiral_xy, xy.data(), sizeof(int16_t) * 2 * c->ncand, cudaMemcpyHostToDevice));
        {
            // zero predictors: centre (0,0) and the rate are the same for every MB, so ((rate + bias) << 15) + key of
            // every candidate is a constant of the context (Gen A cost domain, the packed 32-bit minimum of me_int_tb.cu);
            // the MV (0,0) pre-test of FASTFULL / !rdopt is key 0
            const int lf = c->lambda_factor;
            const unsigned bias = c->p.rdopt ? 0u : (unsigned)(((long long)lf * 16) >> 16);
            const bool pt = !c->p.rdopt && c->p.search_mode == JMME_SEARCH_FASTFULL;
            auto se_bits = [](int v) { int a = v < 0 ? -v : v, n = 1; while (a) { n += 2; a >>= 1; } return n; };
            std::vector<uint32_t> kr((size_t)c->ncand);
            for (int yo = 0; yo < c->ncols; yo++)
                for (int xo = 0; xo < c->ncols; xo++) {
                    const size_t i = (size_t)yo * c->ncols + xo;
                    const unsigned rate = (unsigned)(((long long)lf * (se_bits(4 * (xo - R)) + se_bits(4 * (yo - R)))) >> 16);
                    kr[i] = ((rate + bias) << JMME_KEY_BITS) + ((pt && xo == R && yo == R) ? 0u : (unsigned)key[i]);
                }
            CUC(cudaMalloc(&c->d_kr0, sizeof(uint32_t) * c->ncand));
            CUC(cudaMemcpy(c->d_kr0, kr.data(), sizeof(uint32_t) * c->ncand, cudaMemcpyHostToDevice));
            const size_t gb = sizeof(uint32_t) * JMME_NBLK * n_mb * p->num_refs;
            CUC(cudaMalloc(&c->d_gbest, gb));
            CUC(cudaMemset(c->d_gbest, 0xFF, gb));
            CUC(cudaMalloc(&c->d_ready, sizeof(int) * n_mb * p->num_refs));
            CUC(cudaMemset(c->d_ready, 0, sizeof(int) * n_mb * p->num_refs));
        }
        if (p->pred_policy == JMME_PRED_MEDIAN) {
            // 2:1 wavefront inside every slice: MB (x, y) of a slice that starts at row y0 is decided in step
            // x + 2 (y - y0), after its left, upper-left, upper and upper-right neighbours
            const int rb = c->p.mb_row_begin, re = c->p.mb_row_end;
            const int k = p->slice_rows ? p->slice_rows : c->mb_h;
            std::vector<int> list;
            c->n_steps = c->mb_w + 2 * (std::min(k, re - rb) - 1);
            c->wave_off = (int *)malloc(sizeof(int) * (c->n_steps + 1));
            if (!c->wave_off) { rc = JMME_ERR_NOMEM; goto bad; }
            for (int t = 0; t < c->n_steps; t++) {
                c->wave_off[t] = (int)list.size();
                for (int y = rb; y < re; y++) {
                    const int x = t - 2 * ((y - rb) % k);
                    if (x >= 0 && x < c->mb_w) list.push_back(y * c->mb_w + x);
                }
            }
            c->wave_off[c->n_steps] = (int)list.size();
            CUC(cudaMalloc(&c->d_wave, sizeof(int) * list.size()));
            CUC(cudaMemcpy(c->d_wave, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice));
            CUC(cudaMalloc(&c->d_fmv, sizeof(int16_t) * 2 * 16 * n_mb));
            CUC(cudaMalloc(&c->d_fref, 16 * n_mb));
            CUC(cudaMemset(c->d_fmv, 0, sizeof(int16_t) * 2 * 16 * n_mb));
            CUC(cudaMemset(c->d_fref, 0xFF, 16 * n_mb));
            CUC(cudaMalloc(&c->d_wave_tab, sizeof(WaveTab)));
            CUC(cudaMemcpy(c->d_wave_tab, &jmme_wave_tab(), sizeof(WaveTab), cudaMemcpyHostToDevice));
        }
    }
#undef CUC
    *out = c;
    return JMME_OK;
bad:
    free_device(c);
    delete c;
    return rc;
}

void fill_search_params(const jmme_ctx *c, SearchParams &P, const uint8_t *cur, int cur_stride, const int16_t *d_pred,
                        jmme_mbresult *d_out, jmme_mbresult *d_out_per_ref)
{
    memset(&P, 0, sizeof P);
    P.cur = cur; P.cur_stride = cur_stride;
    for (int r = 0; r < c->p.num_refs; r++) P.planes[r] = c->d_planes[r];
    P.pstride = c->pstride; P.pheight = c->pheight; P.pad = c->pad;
    P.mb_w = c->mb_w; P.mb_h = c->mb_h; P.mb_row_begin = c->p.mb_row_begin; P.mb_row_end = c->p.mb_row_end;
    P.R = c->p.search_range; P.ncols = c->ncols; P.num_refs = c->p.num_refs;
    P.lambda_factor = c->lambda_factor; P.rdopt = c->p.rdopt; P.search_mode = c->p.search_mode; P.cmax = c->cmax;
    P.pred_policy = c->p.pred_policy; P.blocktype_mask = c->p.blocktype_mask;
    P.use_hadamard = c->metric[1] == JMME_DIST_HADAMARD; P.satd_round = c->p.satd_round; P.subpel = c->p.subpel;
    P.pred = c->p.pred_policy == JMME_PRED_ZERO ? nullptr : d_pred;
    P.spiral_key = c->d_spiral_key; P.spiral_xy = c->d_spiral_xy; P.kr0 = c->d_kr0; P.gbest = c->d_gbest;
    P.res = c->d_res; P.out = d_out; P.out_per_ref = d_out_per_ref;
    P.cost_domain = c->p.cost_domain; P.ext = c->ext; P.t8 = c->p.transform8x8; P.chroma_me = c->p.chroma_me;
    for (int st = 0; st < 3; st++) { P.metric[st] = c->metric[st]; P.lf[st] = c->lf[st]; }
    for (int r = 0; r < c->p.num_refs; r++)
        for (int k = 0; k < 2; k++) P.cplanes[r][k] = c->d_cplanes[r][k];
    P.cstride = c->cstride; P.cpad = c->cpad; P.cur_c[0] = c->d_cur_c[0]; P.cur_c[1] = c->d_cur_c[1]; P.cur_cs = c->w16 / 2;
    P.tune_group = c->tune.group; P.tune_cluster = c->tune.cluster; P.tune_lin = !c->tune.table_rate;
    P.tune_split = c->tune.balance != 2;
}

// enqueue the whole search on `st`; nothing is synchronised here
// rb/re: MB-row range to search (a sub-range of the context's stripe; -1 = the whole stripe)
int enqueue_search(jmme_ctx *c, const uint8_t *d_cur, int stride, const int16_t *d_pred, jmme_mbresult *d_out,
                   jmme_mbresult *d_out_per_ref, cudaStream_t st, int rb = -1, int re = -1)
{
    for (int r = 0; r < c->p.num_refs; r++)
        if (!c->ref_set[r]) return fail(c, JMME_ERR_STATE, "reference not set");
    if (c->p.chroma_me) {
        if (!c->cur_c_set) return fail(c, JMME_ERR_STATE, "current chroma not set");
        for (int r = 0; r < c->p.num_refs; r++)
            if (!c->cref_set[r]) return fail(c, JMME_ERR_STATE, "reference chroma not set");
    }
    const uint8_t *cur = d_cur;
    int cs = stride;
    // the search kernel fetches the current MB with 16-byte cp.async: rows must be 16-byte aligned
    // (rows below the picture are handled in the kernels by clamping the row index: only a width that is
    // not a multiple of 16 needs the padded copy)
    int cur_h = c->p.height;
    if (c->w16 != c->p.width || (stride & 15) || ((uintptr_t)d_cur & 15)) {
        CU(c, jmme_launch_pad_cur(d_cur, c->p.width, c->p.height, stride, c->w16, c->h16, c->d_cur16, st));
        c->launches++;
        cur = c->d_cur16; cs = c->w16; cur_h = c->h16;
    }
    SearchParams P;
    fill_search_params(c, P, cur, cs, d_pred, d_out, d_out_per_ref);
    P.cur_h = cur_h;
    P.fused_select = c->p.num_refs == 1 && c->p.subpel;
    if (c->dev_call && c->mc_field) { P.mc_out = (jmme_mbresult *)c->mc_field; P.n_peer_out = -1; }   // multicast gather
    for (int i = 0; c->dev_call && !c->mc_field && i < c->n_peer_fields; i++)          // fused gather (jmme_set_peer_fields_dev)
        if (c->peer_fields[i] && c->peer_fields[i] != (void *)d_out) P.peer_out[P.n_peer_out++] = (jmme_mbresult *)c->peer_fields[i];
    if (rb >= 0) { P.mb_row_begin = rb; P.mb_row_end = re; }
    c->prof_valid[1] = c->prof_valid[2] = c->prof_valid[3] = false;
    if (c->p.pred_policy == JMME_PRED_MEDIAN) {
        // predict -> search -> commit, one wavefront step after the other; the kernels of a step work on
        // that step's list of MBs; the kernel that writes the records of a step — sub-pel or
        // reference selection — commits its MBs to the field
        P.pred = c->d_pred; P.pred_policy = JMME_PRED_PER_BLOCK;
        P.field_mv = c->d_fmv; P.field_ref = c->d_fref; P.slice_rows = c->p.slice_rows;
        // the default integer kernel of a wavefront step (me_int_tb.cu, 12 warps, clusters) predicts in its own
        // prologue; the other kernels read the predictors wave_step_kernel writes
        const bool wide = c->p.cost_domain || c->metric[0] == JMME_DIST_SSE;      // me_int.cu's 64-bit kernels
        const bool in_kernel = c->tune.variant == 0 && !wide && c->p.search_mode == JMME_SEARCH_FASTFULL && c->p.search_range <= 32 &&
                               c->ncols >= 6 && c->p.blocktype_mask != JMME_MASK_16x16 && !c->tune.wave_step;
        P.wave_tab = in_kernel ? c->d_wave_tab : nullptr;
        // search and sub-pel kernels of consecutive steps overlap their launch latency and constant-only
        // prologues (programmatic dependent launch); jmme_tuning.no_pdl turns it off
        P.pdl = in_kernel && !c->tune.no_pdl;
        for (int t = 0; t < c->n_steps; t++) {
            P.mb_list = c->d_wave + c->wave_off[t]; P.n_list = c->wave_off[t + 1] - c->wave_off[t];
            if (!in_kernel) {
                CU(c, jmme_launch_wave_step(P.mb_list, P.n_list, c->mb_w, c->mb_h, c->p.num_refs, c->p.slice_rows,
                                            c->d_fmv, c->d_fref, c->d_pred, st));
                c->launches++;
            }
            if (c->p.search_mode == JMME_SEARCH_FULL) CU(c, jmme_launch_me_full(P, st));
            else CU(c, jmme_launch_me_int(P, c->num_sms, c->tune.variant, st));
            c->launches++;
            if (c->p.subpel) { CU(c, jmme_launch_subpel(P, st)); c->launches++; }
            if (!P.fused_select) { CU(c, jmme_launch_select(P, st)); c->launches++; }
        }
        memcpy(c->last_kernel, jmme_kernel_name_buf(), JMME_KNAME_LEN);
        CU(c, cudaEventRecord(c->ev_search, st));    // jmme_get_predictors copies d_pred on another stream
        c->searched = true;
        return JMME_OK;
    }
    // zero predictors, R = 32: balanced task ranges, packed minima in d_gbest until the next kernel consumes them
    // (me_int_tb.cu BAL); a search that failed half-way may have left words behind
    const bool full_pb = c->p.search_mode == JMME_SEARCH_FULL && c->p.pred_policy == JMME_PRED_PER_BLOCK;
    P.int_packed = !full_pb && jmme_me_int_balanced(P, c->tune.variant, c->num_sms, c->tune.balance == 1);
    if (P.tune_group <= 0) {
        // MBs per item of the zero-predictor kernel, measured: 4 once the launch has two or more rounds of MB pairs
        // on the 3 x SMs resident CTAs (fewer window stagings per MB), else pairs (a short stripe needs the items)
        const long long pairs = (long long)(P.mb_row_end - P.mb_row_begin) * ((P.mb_w + 1) / 2) * P.num_refs;
        P.tune_group = (P.int_packed || pairs >= 6LL * c->num_sms) ? 4 : 2;
        if (P.tune_group == 4 && !P.int_packed && P.num_refs == 1 && !c->tune.no_pair_tail) {
            // whole rounds of 4-MB items on the 3 x SMs resident CTAs, the remaining rows as pairs: the launch ends with a
            // short round of short items instead of a long round of long ones
            const int cap = 3 * c->num_sms, ppr4 = (P.mb_w + 3) / 4, rows = P.mb_row_end - P.mb_row_begin;
            const int rounds = rows * ppr4 / cap;                       // full rounds of 4-MB items available
            const int r4 = std::min(rows, rounds * cap / ppr4);          // rows that fill whole rounds
            if (rounds >= 1 && r4 < rows) P.pair_row = P.mb_row_begin + r4;
        }
    }
    // early start of the sub-pel kernel behind a whole-item search (per-MB ready flags); not while per-kernel event
    // brackets separate the two launches
    P.ready = (c->p.subpel && !full_pb && !c->profiling && c->tune.early_subpel != 2 &&
               jmme_me_int_raises_flags(P, c->tune.variant)) ? c->d_ready : nullptr;
    if (P.ready && c->ready_dirty) {
        CU(c, cudaMemsetAsync(c->d_ready, 0, sizeof(int) * (size_t)c->mb_w * c->mb_h * c->p.num_refs, st));
        c->ready_dirty = false;
    }
    if (P.ready) c->ready_dirty = true;
    if (P.int_packed && c->gbest_dirty) {
        CU(c, cudaMemsetAsync(c->d_gbest, 0xFF, sizeof(uint32_t) * JMME_NBLK * (size_t)c->mb_w * c->mb_h * c->p.num_refs, st));
        c->gbest_dirty = false;
    }
    if (P.int_packed) c->gbest_dirty = true;         // until every launch of this search is queued
    if (c->profiling) CU(c, cudaEventRecord(c->ev_prof[1][0], st));
    if (full_pb)
        CU(c, jmme_launch_me_full(P, st));           // a window per block: nothing to share (me_full.cu)
    else
        CU(c, jmme_launch_me_int(P, c->num_sms, c->tune.variant, st));
    memcpy(c->last_kernel, jmme_kernel_name_buf(), JMME_KNAME_LEN);
    if (c->profiling) { CU(c, cudaEventRecord(c->ev_prof[1][1], st)); c->prof_valid[1] = true; }
    c->launches++;
    if (c->p.subpel) {
        if (c->profiling) CU(c, cudaEventRecord(c->ev_prof[2][0], st));
        CU(c, jmme_launch_subpel(P, st));
        if (c->profiling) { CU(c, cudaEventRecord(c->ev_prof[2][1], st)); c->prof_valid[2] = true; }
        c->launches++;
    }
    if (!P.fused_select) {
        if (c->profiling) CU(c, cudaEventRecord(c->ev_prof[3][0], st));
        CU(c, jmme_launch_select(P, st));
        if (c->profiling) { CU(c, cudaEventRecord(c->ev_prof[3][1], st)); c->prof_valid[3] = true; }
        c->launches++;
    }
    c->gbest_dirty = false;
    c->ready_dirty = false;
    return JMME_OK;
}

}  // namespace

extern "C" {

void jmme_default_params(jmme_params *p)
{
    memset(p, 0, sizeof *p);
    p->search_range = 32; p->num_refs = 1; p->blocktype_mask = JMME_MASK_ALL;
    p->qp = 28; p->use_hadamard = 1;
    p->search_mode = JMME_SEARCH_FASTFULL; p->pred_policy = JMME_PRED_ZERO;
}

int jmme_lambda_factor(int qp, int rdopt)
{
    int q = std::min(std::max(qp - 12, 0), 39);
    double lambda = rdopt ? std::sqrt(0.85 * std::pow(2.0, q / 3.0)) : (double)kQp2Quant[q];
    return (int)(65536.0 * lambda + 0.5);
}

int jmme_create(jmme_ctx **out, const jmme_params *p)
{
    if (!out || !p) return JMME_ERR_PARAM;
    *out = nullptr;
    int rc = validate(p);
    if (rc != JMME_OK) return rc;
    if (p->n_gpus <= 1) return create_single(out, p, p->device_ids[0]);

    // multi-GPU parent: split the stripe's MB rows as evenly as possible over the devices
    const int h16 = (p->height + 15) & ~15, mb_h = h16 / 16;
    const int rb = p->mb_row_begin, re = p->mb_row_end ? p->mb_row_end : mb_h;
    if (rb < 0 || re > mb_h || rb >= re) return JMME_ERR_PARAM;
    jmme_ctx *c = new (std::nothrow) jmme_ctx();
    if (!c) return JMME_ERR_NOMEM;
    memset(c, 0, sizeof *c);
    c->p = *p; c->device = -1;
    // (in-frame median: whole slices per device, so the result does not depend on the device count)
    const int unit = p->pred_policy == JMME_PRED_MEDIAN ? (p->slice_rows ? p->slice_rows : mb_h) : 1;
    if (unit > 1 && (rb % unit || (re % unit && re != mb_h))) { delete c; return JMME_ERR_PARAM; }
    const int rows = (re - rb + unit - 1) / unit, n = std::min(p->n_gpus, rows);
    int r0 = rb;
    for (int g = 0; g < n; g++) {
        jmme_params q = *p;
        q.n_gpus = 1;
        q.mb_row_begin = r0; q.mb_row_end = std::min(re, r0 + unit * (rows / n + (g < rows % n ? 1 : 0)));
        r0 = q.mb_row_end;
        rc = create_single(&c->sub[g], &q, p->device_ids[g]);
        if (rc != JMME_OK) { jmme_destroy(c); return rc; }
        c->n_sub++;
    }
    jmme_ctx *s0 = c->sub[0];
    c->w16 = s0->w16; c->h16 = s0->h16; c->mb_w = s0->mb_w; c->mb_h = s0->mb_h; c->pad = s0->pad; c->max_pred = s0->max_pred; c->cmax = s0->cmax;
    c->pstride = s0->pstride; c->pheight = s0->pheight; c->lambda_factor = s0->lambda_factor;
    c->p.mb_row_begin = rb; c->p.mb_row_end = re;
    for (int g = 1; g < c->n_sub; g++) {         // NVLink peer access towards the gathering device
        int can = 0;
        cudaDeviceCanAccessPeer(&can, s0->device, c->sub[g]->device);
        if (can) {
            cudaSetDevice(s0->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(c->sub[g]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        }
    }
    *out = c;
    return JMME_OK;
}

int jmme_destroy(jmme_ctx *c)
{
    if (!c) return JMME_OK;
    for (int g = 0; g < c->n_sub; g++) jmme_destroy(c->sub[g]);
    if (!c->n_sub && c->device >= 0) free_device(c);
    delete c;
    return JMME_OK;
}

const char *jmme_strerror(int code)
{
    switch (code) {
    case JMME_OK: return "ok";
    case JMME_ERR_PARAM: return "invalid parameter";
    case JMME_ERR_CUDA: return "CUDA error";
    case JMME_ERR_NOMEM: return "out of memory";
    case JMME_ERR_UNSUPPORTED: return "unsupported configuration";
    case JMME_ERR_STATE: return "invalid state (reference not set?)";
    case JMME_ERR_NODEVICE: return "no CUDA device";
    default: return "unknown error";
    }
}
const char *jmme_last_error(const jmme_ctx *c) { return c ? c->err : ""; }
const char *jmme_backend(void) { return "cuda-sm_100a"; }
int jmme_abi_version(void) { return JMME_ABI_VERSION; }
int jmme_mb_width(const jmme_ctx *c) { return c ? c->mb_w : 0; }
int jmme_mb_height(const jmme_ctx *c) { return c ? c->mb_h : 0; }
int jmme_pad(const jmme_ctx *c) { return c ? c->pad : 0; }
int jmme_lambda_factor_of(const jmme_ctx *c) { return c ? c->lambda_factor : 0; }
int64_t jmme_launch_count(const jmme_ctx *c)
{
    if (!c) return 0;
    long long n = c->launches;
    for (int g = 0; g < c->n_sub; g++) n += c->sub[g]->launches;
    return n;
}

static int flush_pending_refs(jmme_ctx *c);
int jmme_set_tuning(jmme_ctx *c, const jmme_tuning *t)
{
    if (!c || !t) return JMME_ERR_PARAM;
    if (t->variant < 0 || t->group < 0 || t->group > 4 || t->cluster < 0 || t->cluster > 4 || t->pipe_parts < 0)
        return fail(c, JMME_ERR_PARAM, "tuning out of range");
    if (!c->n_sub) {                              // reference chunks still pending were cut for the old pipe_parts
        if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, JMME_ERR_CUDA, "cudaSetDevice");
        int rcf = flush_pending_refs(c);
        if (rcf != JMME_OK) return rcf;
    }
    resolve_tuning(t, &c->tune);
    for (int g = 0; g < c->n_sub; g++) resolve_tuning(t, &c->sub[g]->tune);
    return JMME_OK;
}
int jmme_get_tuning(const jmme_ctx *c, jmme_tuning *t)
{
    if (!c || !t) return JMME_ERR_PARAM;
    *t = c->n_sub ? c->sub[0]->tune : c->tune;
    return JMME_OK;
}
const char *jmme_last_kernel(const jmme_ctx *c) { return !c ? "" : (c->n_sub ? c->sub[0]->last_kernel : c->last_kernel); }

int jmme_set_reference_dev(jmme_ctx *c, int r, const void *d_luma, int stride, void *stream)
{
    if (!c || !d_luma || r < 0 || r >= c->p.num_refs || stride < c->p.width) return JMME_ERR_PARAM;
    if (c->n_sub) return fail(c, JMME_ERR_UNSUPPORTED, "device-pointer calls need a single-device context");
    CU(c, cudaSetDevice(c->device));
    // only the plane rows this context's stripe can reach: window centre within +-R, window +-R,
    // 16 rows of the MB, one more sample for the sub-pel candidates (rounded out to 4)
    const int R = c->p.search_range;
    // (a stripe that touches the top / bottom of the picture produces the whole border)
    const int y_begin = c->p.mb_row_begin == 0 ? 0 : std::max(0, c->pad + 16 * c->p.mb_row_begin - c->cmax - R - 4);
    const int y_end = c->p.mb_row_end == c->mb_h ? c->pheight
                                                 : std::min(c->pheight, c->pad + 16 * c->p.mb_row_end + c->cmax + R + 4);
    c->prof_valid[0] = false;
    if (c->profiling) CU(c, cudaEventRecord(c->ev_prof[0][0], (cudaStream_t)stream));
    CU(c, jmme_launch_interp((const uint8_t *)d_luma, c->p.width, c->p.height, stride, c->pad, c->pstride, c->pheight,
                             c->n_planes, c->d_planes[r], y_begin, y_end, (cudaStream_t)stream));
    if (c->profiling) { CU(c, cudaEventRecord(c->ev_prof[0][1], (cudaStream_t)stream)); c->prof_valid[0] = true; }
    c->launches++;
    c->ref_set[r] = true;
    return JMME_OK;
}

// First MB row of part k of np of the pipelined host path (k = np: the end).  The first part is small, so that the
// first search kernel waits for few rows of both pictures, and the last one is smaller than the middle ones,
// because its download is not hidden behind any kernel.
static int part_begin(const jmme_ctx *c, int k, int np)
{
    static const int kCut[5][5] = {{0, 0, 0, 0, 0}, {0, 100, 0, 0, 0}, {0, 40, 100, 0, 0}, {0, 20, 64, 100, 0}, {0, 12, 42, 75, 100}};
    const int rows = c->p.mb_row_end - c->p.mb_row_begin;
    if (k <= 0) return c->p.mb_row_begin;
    if (k >= np) return c->p.mb_row_end;
    const int cut = c->tune.even_parts ? 100 * k / np : kCut[np][k];
    return c->p.mb_row_begin + std::min(std::max(rows * cut / 100, k), rows - (np - k));
}

// ---- reference upload in chunks (host path) -----------------------------------------------------------------
// Can jmme_search_frame take its pipelined form for this context?  (The properties of the current-picture buffer
// are checked there; if they fail, the pending chunks are flushed first.)
static bool pipelined_ctx(const jmme_ctx *c)
{
    return !c->n_sub && !c->profiling && c->tune.pipe_parts > 1 && c->p.pred_policy != JMME_PRED_MEDIAN &&
           c->w16 == c->p.width && c->p.mb_row_end - c->p.mb_row_begin >= 4 * c->tune.pipe_parts;
}
// plane rows [*pb, *pe) and picture rows [*sb, *se) of chunk k of np: chunk k ends where part k of the search can
// reach (its MB rows + window + centre excursion + sub-pel margin); a plane row needs picture rows -2 .. +3
static void ref_chunk(const jmme_ctx *c, int k, int np, int *pb, int *pe, int *sb, int *se)
{
    const int R = c->p.search_range;
    const int yb = c->p.mb_row_begin == 0 ? 0 : std::max(0, c->pad + 16 * c->p.mb_row_begin - c->cmax - R - 4);
    const int ye = c->p.mb_row_end == c->mb_h ? c->pheight : std::min(c->pheight, c->pad + 16 * c->p.mb_row_end + c->cmax + R + 4);
    auto end_of = [&](int j) {
        if (j < 0) return yb;
        if (j >= np - 1) return ye;
        const int re = part_begin(c, j + 1, np);
        return std::min(ye, c->pad + 16 * re + c->cmax + R + 4);
    };
    auto src_end = [&](int y) { return std::min(std::max(y - c->pad + 3, 1), c->p.height); };     // exclusive
    *pb = end_of(k - 1); *pe = end_of(k);
    *sb = k == 0 ? std::min(std::max(yb - c->pad - 3, 0), c->p.height - 1) : src_end(*pb);
    *se = src_end(*pe);
}
// queue chunks [from, to) of reference r on the context's stream: picture rows up, their plane rows interpolated
static int queue_ref_chunks(jmme_ctx *c, int r, const uint8_t *luma, int stride, int from, int to, int np)
{
    for (int k = from; k < to; k++) {
        int pb, pe, sb, se;
        ref_chunk(c, k, np, &pb, &pe, &sb, &se);
        if (se > sb)
            CU(c, upload_rows(c->d_raw_ref[r] + (size_t)sb * c->p.width, luma + (size_t)sb * stride, stride, c->p.width, se - sb,
                              c->stream));
        if (pe > pb) {
            CU(c, jmme_launch_interp(c->d_raw_ref[r], c->p.width, c->p.height, c->p.width, c->pad, c->pstride, c->pheight,
                                     c->n_planes, c->d_planes[r], pb, pe, c->stream));
            c->launches++;
        }
    }
    return JMME_OK;
}
// everything still pending goes out now (a call that reads the planes outside the pipelined search)
static int flush_pending_refs(jmme_ctx *c)
{
    for (int r = 0; r < c->p.num_refs; r++)
        if (c->pend_next[r]) {
            int rc = queue_ref_chunks(c, r, c->pend_luma[r], c->pend_stride[r], c->pend_next[r], c->tune.pipe_parts, c->tune.pipe_parts);
            c->pend_next[r] = 0;
            if (rc != JMME_OK) return rc;
        }
    return JMME_OK;
}

int jmme_set_reference(jmme_ctx *c, int r, const uint8_t *luma, int stride)
{
    if (!c || !luma || r < 0 || r >= c->p.num_refs || stride < c->p.width) return JMME_ERR_PARAM;
    if (c->n_sub) {
        for (int g = 0; g < c->n_sub; g++) {
            int rc = jmme_set_reference(c->sub[g], r, luma, stride);
            if (rc != JMME_OK) return fail(c, rc, c->sub[g]->err);
        }
        return JMME_OK;
    }
    CU(c, cudaSetDevice(c->device));
    c->pend_next[r] = 0;
    if (c->p.async_reference && pipelined_ctx(c)) {
        // first chunk now, the others with the parts of the next jmme_search_frame (the caller keeps `luma` unchanged
        // until that call returns: the contract of async_reference)
        int rc = queue_ref_chunks(c, r, luma, stride, 0, 1, c->tune.pipe_parts);
        if (rc != JMME_OK) return rc;
        c->pend_luma[r] = luma; c->pend_stride[r] = stride; c->pend_next[r] = 1;
        c->ref_set[r] = true;
        return JMME_OK;
    }
    {
        // upload only the picture rows the stripe's planes are built from (plane rows +-3 filter taps)
        const int R = c->p.search_range;
        const int yb = c->p.mb_row_begin == 0 ? 0 : std::max(0, c->pad + 16 * c->p.mb_row_begin - c->cmax - R - 4);
        const int ye = c->p.mb_row_end == c->mb_h ? c->pheight : std::min(c->pheight, c->pad + 16 * c->p.mb_row_end + c->cmax + R + 4);
        const int s0 = std::min(std::max(yb - c->pad - 3, 0), c->p.height - 1);
        const int s1 = std::min(std::max(ye - c->pad + 3, 1), c->p.height);          // exclusive; clamps hit row h-1
        CU(c, upload_rows(c->d_raw_ref[r] + (size_t)s0 * c->p.width, luma + (size_t)s0 * stride, stride, c->p.width, s1 - s0,
                          c->stream));
    }
    int rc = jmme_set_reference_dev(c, r, c->d_raw_ref[r], c->p.width, c->stream);
    if (rc != JMME_OK) return rc;
    if (!c->p.async_reference) CU(c, cudaStreamSynchronize(c->stream));
    return JMME_OK;
}

// chroma ME: replicate a chroma pair (device pointers) into the padded planes of reference r (r >= 0) or into the
// current-chroma buffers (r < 0), asynchronously on `st`
static int chroma_to_planes(jmme_ctx *c, int r, const uint8_t *d_cb, const uint8_t *d_cr, int stride, cudaStream_t st)
{
    const int cw = (c->p.width + 1) / 2, ch = (c->p.height + 1) / 2;
    const uint8_t *src[2] = {d_cb, d_cr};
    for (int k = 0; k < 2; k++) {
        if (r >= 0)
            CU(c, jmme_launch_pad_plane(src[k], cw, ch, stride, c->cpad, c->cstride, c->cheight, c->d_cplanes[r][k], st));
        else
            CU(c, jmme_launch_pad_plane(src[k], cw, ch, stride, 0, c->w16 / 2, c->h16 / 2, c->d_cur_c[k], st));
        c->launches++;
    }
    if (r >= 0) c->cref_set[r] = true;
    else c->cur_c_set = true;
    return JMME_OK;
}

int jmme_set_reference_chroma_dev(jmme_ctx *c, int r, const void *d_cb, const void *d_cr, int stride, void *stream)
{
    if (!c || !d_cb || !d_cr || r < 0 || r >= c->p.num_refs || stride < (c->p.width + 1) / 2) return JMME_ERR_PARAM;
    if (c->n_sub) return fail(c, JMME_ERR_UNSUPPORTED, "device-pointer calls need a single-device context");
    if (!c->p.chroma_me) return fail(c, JMME_ERR_STATE, "chroma_me is off");
    CU(c, cudaSetDevice(c->device));
    return chroma_to_planes(c, r, (const uint8_t *)d_cb, (const uint8_t *)d_cr, stride, (cudaStream_t)stream);
}

int jmme_set_current_chroma_dev(jmme_ctx *c, const void *d_cb, const void *d_cr, int stride, void *stream)
{
    if (!c || !d_cb || !d_cr || stride < (c->p.width + 1) / 2) return JMME_ERR_PARAM;
    if (c->n_sub) return fail(c, JMME_ERR_UNSUPPORTED, "device-pointer calls need a single-device context");
    if (!c->p.chroma_me) return fail(c, JMME_ERR_STATE, "chroma_me is off");
    CU(c, cudaSetDevice(c->device));
    return chroma_to_planes(c, -1, (const uint8_t *)d_cb, (const uint8_t *)d_cr, stride, (cudaStream_t)stream);
}

// host chroma: upload both components (whole pictures: a quarter of the luma each), then replicate
static int chroma_from_host(jmme_ctx *c, int r, const uint8_t *cb, const uint8_t *cr, int stride)
{
    if (c->n_sub) {
        for (int g = 0; g < c->n_sub; g++) {
            int rc = chroma_from_host(c->sub[g], r, cb, cr, stride);
            if (rc != JMME_OK) return fail(c, rc, c->sub[g]->err);
        }
        return JMME_OK;
    }
    if (!c->p.chroma_me) return fail(c, JMME_ERR_STATE, "chroma_me is off");
    const int cw = (c->p.width + 1) / 2, ch = (c->p.height + 1) / 2;
    CU(c, cudaSetDevice(c->device));
    // the staging buffer is reused by the next call: the stream is drained first (chroma ME is not the pipelined path)
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, upload_rows(c->d_craw, cb, stride, cw, ch, c->stream));
    CU(c, upload_rows(c->d_craw + (size_t)cw * ch, cr, stride, cw, ch, c->stream));
    int rc = chroma_to_planes(c, r, c->d_craw, c->d_craw + (size_t)cw * ch, cw, c->stream);
    if (rc != JMME_OK) return rc;
    CU(c, cudaStreamSynchronize(c->stream));
    return JMME_OK;
}

int jmme_set_reference_chroma(jmme_ctx *c, int r, const uint8_t *cb, const uint8_t *cr, int stride)
{
    if (!c || !cb || !cr || r < 0 || r >= c->p.num_refs || stride < (c->p.width + 1) / 2) return JMME_ERR_PARAM;
    return chroma_from_host(c, r, cb, cr, stride);
}

int jmme_set_current_chroma(jmme_ctx *c, const uint8_t *cb, const uint8_t *cr, int stride)
{
    if (!c || !cb || !cr || stride < (c->p.width + 1) / 2) return JMME_ERR_PARAM;
    return chroma_from_host(c, -1, cb, cr, stride);
}

// ---- (f2) bi-predictive refinement ----------------------------------------------------------------------------
int jmme_set_reference_l1(jmme_ctx *c, const uint8_t *luma, int stride)
{
    if (!c || !luma || stride < c->p.width) return JMME_ERR_PARAM;
    if (c->n_sub) {
        for (int g = 0; g < c->n_sub; g++) {
            int rc = jmme_set_reference_l1(c->sub[g], luma, stride);
            if (rc != JMME_OK) return fail(c, rc, c->sub[g]->err);
        }
        return JMME_OK;
    }
    CU(c, cudaSetDevice(c->device));
    const size_t n_mb = (size_t)c->mb_w * c->mb_h, psz = (size_t)c->pstride * c->pheight;
    if (!c->d_planes_l1) {
        const size_t n_pred = sizeof(int16_t) * 2 * JMME_NBLK * n_mb;
        CU(c, cudaMalloc(&c->d_planes_l1, psz * c->n_planes));
        CU(c, cudaMalloc(&c->d_raw_l1, (size_t)c->p.width * c->p.height));
        CU(c, cudaMalloc(&c->d_bi_l0, sizeof(jmme_mbresult) * n_mb));
        CU(c, cudaMalloc(&c->d_bi_l1, sizeof(jmme_mbresult) * n_mb));
        CU(c, cudaMalloc(&c->d_bi_pred0, n_pred * c->p.num_refs));
        CU(c, cudaMalloc(&c->d_bi_pred1, n_pred));
        CU(c, cudaMalloc(&c->d_bi_spiral, sizeof(int16_t) * 2 * 31 * 31));
        CU(c, cudaMalloc(&c->d_bi_out, sizeof(jmme_bipred) * n_mb));
        CU(c, cudaMalloc(&c->d_bi_err, sizeof(int)));
    }
    CU(c, upload_rows(c->d_raw_l1, luma, stride, c->p.width, c->p.height, c->stream));
    // every plane row: the refined vectors may leave the rows a uni-directional search of this stripe can reach
    CU(c, jmme_launch_interp(c->d_raw_l1, c->p.width, c->p.height, c->p.width, c->pad, c->pstride, c->pheight, c->n_planes,
                             c->d_planes_l1, 0, c->pheight, c->stream));
    c->launches++;
    CU(c, cudaStreamSynchronize(c->stream));
    c->l1_set = true;
    return JMME_OK;
}

int jmme_search_frame_bipred(jmme_ctx *c, const uint8_t *cur, int stride, const jmme_mbresult *l0, const jmme_mbresult *l1,
                             const int16_t *pred0, const int16_t *pred1, int range, int iterations, jmme_bipred *out)
{
    if (!c || !cur || !l0 || !l1 || !out || stride < c->p.width || range < 1 || range > 15 || iterations < 1 || iterations > 8)
        return JMME_ERR_PARAM;
    if (c->n_sub) {
        for (int g = 0; g < c->n_sub; g++) {
            int rc = jmme_search_frame_bipred(c->sub[g], cur, stride, l0, l1, pred0, pred1, range, iterations, out);
            if (rc != JMME_OK) return fail(c, rc, c->sub[g]->err);
        }
        return JMME_OK;
    }
    if (!c->l1_set) return fail(c, JMME_ERR_STATE, "list-1 reference not set");
    { int rcf = flush_pending_refs(c); if (rcf != JMME_OK) return rcf; }
    for (int r = 0; r < c->p.num_refs; r++)
        if (!c->ref_set[r]) return fail(c, JMME_ERR_STATE, "reference not set");
    // list-0 planes of a stripe context cover only the rows its own search reaches: the refinement needs the same
    // margin the header promises (pad - 1), so a stripe context must hold whole planes
    if (c->p.mb_row_begin != 0 || c->p.mb_row_end != c->mb_h) {
        // rows outside [y_begin, y_end) of the list-0 planes were never written: refuse instead of reading them
        const int R = c->p.search_range;
        // quarter-pel, worst case: a uni-directional vector reaches R around the window centre, the centre R around
        // (0,0) when predictors are given; each iteration of a list moves it by up to `range`
        const int reach = 4 * (c->p.pred_policy == JMME_PRED_ZERO ? R : c->cmax + R) + 3 + 4 * range * ((iterations + 1) / 2);
        const int yb = c->p.mb_row_begin == 0 ? 0 : std::max(0, c->pad + 16 * c->p.mb_row_begin - c->cmax - R - 4);
        const int ye = c->p.mb_row_end == c->mb_h ? c->pheight : std::min(c->pheight, c->pad + 16 * c->p.mb_row_end + c->cmax + R + 4);
        const int need_b = c->pad + 16 * c->p.mb_row_begin - std::min((reach >> 2) + 1, c->pad - 1);
        const int need_e = c->pad + 16 * c->p.mb_row_end + std::min((reach >> 2) + 1, c->pad - 1);
        if (need_b < yb || need_e > ye) return fail(c, JMME_ERR_UNSUPPORTED, "bi-pred refinement on a stripe context: range * iterations exceeds the halo");
    }
    CU(c, cudaSetDevice(c->device));
    const size_t n_mb = (size_t)c->mb_w * c->mb_h;
    const int npb = c->p.pred_policy >= JMME_PRED_PER_BLOCK ? JMME_NBLK : 1;
    const size_t off = (size_t)c->p.mb_row_begin * c->mb_w, cnt = (size_t)(c->p.mb_row_end - c->p.mb_row_begin) * c->mb_w;
    for (int k = 0; k < 2; k++) {                           // predictor range, the stripe's rows only
        const int16_t *pr = k ? pred1 : pred0;
        for (int r = 0; pr && r < (k ? 1 : c->p.num_refs); r++) {
            const int16_t *q = pr + ((size_t)r * n_mb + off) * npb * 2;
            for (size_t i = 0; i < cnt * npb * 2; i++)
                if (q[i] > c->max_pred || q[i] < -c->max_pred) return fail(c, JMME_ERR_PARAM, "pred out of range");
        }
    }
    // spiral of the refinement range
    {
        const int nc = (2 * range + 1) * (2 * range + 1);
        std::vector<int16_t> xy(2 * (size_t)nc);
        for (int dy = -range; dy <= range; dy++)
            for (int dx = -range; dx <= range; dx++) {
                const int k = spiral_index(dx, dy);
                xy[2 * (size_t)k] = (int16_t)dx; xy[2 * (size_t)k + 1] = (int16_t)dy;
            }
        CU(c, cudaMemcpyAsync(c->d_bi_spiral, xy.data(), sizeof(int16_t) * 2 * nc, cudaMemcpyHostToDevice, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));            // xy leaves scope
    }
    {
        const int s0 = std::min(16 * c->p.mb_row_begin, c->p.height - 1), s1 = std::min(16 * c->p.mb_row_end, c->p.height);
        CU(c, upload_rows(c->d_raw + (size_t)s0 * c->p.width, cur + (size_t)s0 * stride, stride, c->p.width, std::max(s1 - s0, 1), c->stream));
    }
    CU(c, cudaMemcpyAsync(c->d_bi_l0 + off, l0 + off, cnt * sizeof(jmme_mbresult), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->d_bi_l1 + off, l1 + off, cnt * sizeof(jmme_mbresult), cudaMemcpyHostToDevice, c->stream));
    if (pred0)
        for (int r = 0; r < c->p.num_refs; r++)
            CU(c, cudaMemcpyAsync(c->d_bi_pred0 + ((size_t)r * n_mb + off) * npb * 2, pred0 + ((size_t)r * n_mb + off) * npb * 2,
                                  cnt * npb * 2 * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream));
    if (pred1)
        CU(c, cudaMemcpyAsync(c->d_bi_pred1 + off * npb * 2, pred1 + off * npb * 2, cnt * npb * 2 * sizeof(int16_t),
                              cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemsetAsync(c->d_bi_err, 0, sizeof(int), c->stream));
    const uint8_t *d_cur = c->d_raw;
    int cs = c->p.width, cur_h = c->p.height;
    if (c->w16 != c->p.width) {
        CU(c, jmme_launch_pad_cur(c->d_raw, c->p.width, c->p.height, c->p.width, c->w16, c->h16, c->d_cur16, c->stream));
        c->launches++;
        d_cur = c->d_cur16; cs = c->w16; cur_h = c->h16;
    }
    SearchParams P;
    fill_search_params(c, P, d_cur, cs, nullptr, nullptr, nullptr);
    P.cur_h = cur_h;
    BipredArgs A;
    A.planes_l1 = c->d_planes_l1; A.l0 = c->d_bi_l0; A.l1 = c->d_bi_l1;
    A.pred0 = pred0 ? c->d_bi_pred0 : nullptr; A.pred1 = pred1 ? c->d_bi_pred1 : nullptr;
    A.spiral_xy = c->d_bi_spiral; A.range = range; A.iterations = iterations; A.npb = npb; A.n_planes = c->n_planes;
    A.out = c->d_bi_out; A.err = c->d_bi_err;
    CU(c, jmme_launch_bipred(P, A, c->stream));
    c->launches++;
    int err = 0;
    CU(c, cudaMemcpyAsync(out + off, c->d_bi_out + off, cnt * sizeof(jmme_bipred), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(&err, c->d_bi_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return err ? fail(c, JMME_ERR_PARAM, "l0 / l1 records: reference index or vector phase not usable") : JMME_OK;
}

int jmme_set_profiling(jmme_ctx *c, int enable)
{
    if (!c) return JMME_ERR_PARAM;
    if (c->n_sub) return fail(c, JMME_ERR_UNSUPPORTED, "profiling needs a single-device context");
    CU(c, cudaSetDevice(c->device));
    if (enable)
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 2; j++)
                if (!c->ev_prof[i][j]) CU(c, cudaEventCreate(&c->ev_prof[i][j]));
    c->profiling = enable != 0;
    return JMME_OK;
}

int jmme_get_kernel_times(jmme_ctx *c, float ms[4])
{
    if (!c || !ms) return JMME_ERR_PARAM;
    if (c->n_sub || !c->profiling) return fail(c, JMME_ERR_STATE, "profiling is not enabled");
    CU(c, cudaSetDevice(c->device));
    for (int i = 0; i < 4; i++) {
        ms[i] = 0.f;
        if (!c->prof_valid[i]) continue;
        CU(c, cudaEventSynchronize(c->ev_prof[i][1]));
        CU(c, cudaEventElapsedTime(&ms[i], c->ev_prof[i][0], c->ev_prof[i][1]));
    }
    return JMME_OK;
}

int jmme_get_predictors(jmme_ctx *c, int16_t *pred)
{
    if (!c || !pred) return JMME_ERR_PARAM;
    if (c->p.pred_policy != JMME_PRED_MEDIAN) return fail(c, JMME_ERR_STATE, "no median search yet");
    jmme_ctx *subs1[1] = {c};
    jmme_ctx **subs = c->n_sub ? c->sub : subs1;
    const size_t n_mb = (size_t)c->mb_w * c->mb_h, per_mb = 2 * JMME_NBLK;
    for (int g = 0; g < (c->n_sub ? c->n_sub : 1); g++) {         // every device holds its own stripe's predictors
        jmme_ctx *s = subs[g];
        if (!s->searched) return fail(c, JMME_ERR_STATE, "no median search yet");
        CU(c, cudaSetDevice(s->device));
        CU(c, cudaStreamWaitEvent(s->stream, s->ev_search, 0));     // the search may have run on the caller's stream
        const size_t off = (size_t)s->p.mb_row_begin * s->mb_w, cnt = (size_t)(s->p.mb_row_end - s->p.mb_row_begin) * s->mb_w;
        for (int r = 0; r < c->p.num_refs; r++)
            CU(c, cudaMemcpyAsync(pred + (r * n_mb + off) * per_mb, s->d_pred + (r * n_mb + off) * per_mb,
                                  cnt * per_mb * sizeof(int16_t), cudaMemcpyDeviceToHost, s->stream));
        CU(c, cudaStreamSynchronize(s->stream));
    }
    return JMME_OK;
}

int jmme_get_subimage(jmme_ctx *c, int r, int xf, int yf, uint8_t *dst, int dst_stride)
{
    if (!c || !dst || r < 0 || r >= c->p.num_refs || xf < 0 || xf > 3 || yf < 0 || yf > 3 || dst_stride < c->pstride)
        return JMME_ERR_PARAM;
    if (c->n_sub) return jmme_get_subimage(c->sub[0], r, xf, yf, dst, dst_stride);
    if (!c->ref_set[r]) return JMME_ERR_STATE;
    if ((xf || yf) && c->n_planes != 16) return JMME_ERR_STATE;
    CU(c, cudaSetDevice(c->device));
    { int rc = flush_pending_refs(c); if (rc != JMME_OK) return rc; }
    const uint8_t *src = c->d_planes[r] + (size_t)c->pstride * c->pheight * (yf * 4 + xf);
    CU(c, cudaMemcpy2DAsync(dst, dst_stride, src, c->pstride, c->pstride, c->pheight, cudaMemcpyDeviceToHost,
                            c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return JMME_OK;
}

int jmme_search_frame_dev(jmme_ctx *c, const void *d_cur, int stride, const void *d_pred, void *d_out,
                          void *d_out_per_ref, void *stream)
{
    if (!c || !d_cur || !d_out || stride < c->p.width) return JMME_ERR_PARAM;
    if (c->n_sub) return fail(c, JMME_ERR_UNSUPPORTED, "device-pointer calls need a single-device context");
    if (c->p.pred_policy != JMME_PRED_ZERO && c->p.pred_policy != JMME_PRED_MEDIAN && !d_pred)
        return fail(c, JMME_ERR_PARAM, "pred required");
    CU(c, cudaSetDevice(c->device));
    {
        bool pending = false;                                 // a host-path reference still in chunks (async_reference)
        for (int r = 0; r < c->p.num_refs; r++) pending |= c->pend_next[r] != 0;
        int rcf = flush_pending_refs(c);
        if (rcf != JMME_OK) return rcf;
        if (pending) {                                        // its planes are built on the internal stream
            CU(c, cudaEventRecord(c->ev_ref, c->stream));
            CU(c, cudaStreamWaitEvent((cudaStream_t)stream, c->ev_ref, 0));
        }
    }
    c->dev_call = true;
    const int rc = enqueue_search(c, (const uint8_t *)d_cur, stride, (const int16_t *)d_pred, (jmme_mbresult *)d_out,
                                  (jmme_mbresult *)d_out_per_ref, (cudaStream_t)stream);
    c->dev_call = false;
    return rc;
}

int jmme_commit_field(jmme_ctx *c, const jmme_mbresult *res, int16_t *mv4, int8_t *ref4, uint8_t *mode)
{
    if (!c || !res || !mv4 || !ref4 || !mode) return JMME_ERR_PARAM;
    jmme_ctx *s = c->n_sub ? c->sub[0] : c;
    const size_t n_mb = (size_t)s->mb_w * s->mb_h, cells = 16 * n_mb;
    CU(c, cudaSetDevice(s->device));
    int16_t *d_mv = nullptr; int8_t *d_ref = nullptr; uint8_t *d_mode = nullptr;
    int rc = JMME_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&d_mv, cells * 4)) != cudaSuccess || (e = cudaMalloc(&d_ref, cells)) != cudaSuccess ||
        (e = cudaMalloc(&d_mode, 5 * n_mb)) != cudaSuccess)
        rc = fail(c, JMME_ERR_CUDA, "cudaMalloc(field)", e);
    if (rc == JMME_OK && (e = cudaMemcpyAsync(s->d_out, res, n_mb * sizeof(jmme_mbresult), cudaMemcpyHostToDevice, s->stream)) != cudaSuccess)
        rc = fail(c, JMME_ERR_CUDA, "cudaMemcpyAsync(res)", e);
    if (rc == JMME_OK && (e = jmme_launch_commit(s->d_out, s->mb_w, s->mb_h, s->p.blocktype_mask, d_mv, d_ref, d_mode, s->stream)) != cudaSuccess)
        rc = fail(c, JMME_ERR_CUDA, "commit_kernel", e);
    if (rc == JMME_OK) {
        s->launches++;
        cudaMemcpyAsync(mv4, d_mv, cells * 4, cudaMemcpyDeviceToHost, s->stream);
        cudaMemcpyAsync(ref4, d_ref, cells, cudaMemcpyDeviceToHost, s->stream);
        cudaMemcpyAsync(mode, d_mode, 5 * n_mb, cudaMemcpyDeviceToHost, s->stream);
        if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess) rc = fail(c, JMME_ERR_CUDA, "commit_field", e);
    }
    cudaFree(d_mv); cudaFree(d_ref); cudaFree(d_mode);
    return rc;
}

int jmme_predict_frame(jmme_ctx *c, const int16_t *mv4, const int8_t *ref4, int16_t *pred)
{
    if (!c || !mv4 || !ref4 || !pred) return JMME_ERR_PARAM;
    jmme_ctx *s = c->n_sub ? c->sub[0] : c;
    const size_t n_mb = (size_t)s->mb_w * s->mb_h, cells = 16 * n_mb;
    const size_t n_pred = (size_t)s->p.num_refs * n_mb * JMME_NBLK * 2;
    CU(c, cudaSetDevice(s->device));
    int16_t *d_mv = nullptr, *d_p = nullptr; int8_t *d_ref = nullptr;
    int rc = JMME_OK;
    cudaError_t e;
    // (a scratch buffer of its own: d_pred keeps the predictors of the last in-frame median search)
    if ((e = cudaMalloc(&d_mv, cells * 4)) != cudaSuccess || (e = cudaMalloc(&d_ref, cells)) != cudaSuccess ||
        (e = cudaMalloc(&d_p, n_pred * sizeof(int16_t))) != cudaSuccess)
        rc = fail(c, JMME_ERR_CUDA, "cudaMalloc(field)", e);
    if (rc == JMME_OK) {
        cudaMemcpyAsync(d_mv, mv4, cells * 4, cudaMemcpyHostToDevice, s->stream);
        cudaMemcpyAsync(d_ref, ref4, cells, cudaMemcpyHostToDevice, s->stream);
        if ((e = jmme_launch_predict(d_mv, d_ref, s->mb_w, s->mb_h, s->p.num_refs, d_p, s->stream)) != cudaSuccess)
            rc = fail(c, JMME_ERR_CUDA, "predict_kernel", e);
    }
    if (rc == JMME_OK) {
        s->launches++;
        cudaMemcpyAsync(pred, d_p, n_pred * sizeof(int16_t), cudaMemcpyDeviceToHost, s->stream);
        if ((e = cudaStreamSynchronize(s->stream)) != cudaSuccess) rc = fail(c, JMME_ERR_CUDA, "predict_frame", e);
    }
    cudaFree(d_mv); cudaFree(d_ref); cudaFree(d_p);
    return rc;
}

int jmme_set_peer_fields_dev(jmme_ctx *c, void *const *d_peers, int n_peers)
{
    if (!c || n_peers < 0 || n_peers > JMME_MAX_GPUS || (n_peers && !d_peers)) return JMME_ERR_PARAM;
    if (c->n_sub) return fail(c, JMME_ERR_UNSUPPORTED, "device-pointer calls need a single-device context");
    for (int i = 0; i < n_peers; i++) c->peer_fields[i] = d_peers[i];
    c->n_peer_fields = n_peers;
    return JMME_OK;
}

int jmme_set_multicast_field_dev(jmme_ctx *c, void *d_field_multicast)
{
    if (!c) return JMME_ERR_PARAM;
    if (c->n_sub) return fail(c, JMME_ERR_UNSUPPORTED, "device-pointer calls need a single-device context");
    c->mc_field = d_field_multicast;
    return JMME_OK;
}

int jmme_push_stripe_dev(jmme_ctx *c, const void *d_local, void *const *d_peers, int n_peers, void *stream)
{
    if (!c || !d_local || !d_peers || n_peers < 1 || n_peers > JMME_MAX_GPUS) return JMME_ERR_PARAM;
    if (c->n_sub) return fail(c, JMME_ERR_UNSUPPORTED, "device-pointer calls need a single-device context");
    CU(c, cudaSetDevice(c->device));
    const size_t off = (size_t)c->p.mb_row_begin * c->mb_w * sizeof(jmme_mbresult);          // bytes, multiple of 4
    const size_t n_words = (size_t)(c->p.mb_row_end - c->p.mb_row_begin) * c->mb_w * sizeof(jmme_mbresult) / 4;
    uint32_t *dst[JMME_MAX_GPUS];
    int n = 0;
    for (int i = 0; i < n_peers; i++)
        if (d_peers[i] && d_peers[i] != d_local) dst[n++] = (uint32_t *)((uint8_t *)d_peers[i] + off);
    if (!n) return JMME_OK;
    CU(c, jmme_launch_push((const uint32_t *)((const uint8_t *)d_local + off), dst, n, n_words, (cudaStream_t)stream));
    c->launches++;
    return JMME_OK;
}

int jmme_search_frame(jmme_ctx *c, const uint8_t *cur, int stride, const int16_t *pred, jmme_mbresult *out,
                      jmme_mbresult *out_per_ref)
{
    if (!c || !cur || !out || stride < c->p.width) return JMME_ERR_PARAM;
    if (c->p.pred_policy == JMME_PRED_MEDIAN) pred = nullptr;
    else if (c->p.pred_policy != JMME_PRED_ZERO && !pred) return fail(c, JMME_ERR_PARAM, "pred required");
    jmme_ctx *subs1[1] = {c};
    jmme_ctx **subs = c->n_sub ? c->sub : subs1;
    const int ns = c->n_sub ? c->n_sub : 1;
    const size_t n_mb = (size_t)c->mb_w * c->mb_h;
    const int npb = c->p.pred_policy == JMME_PRED_PER_BLOCK ? JMME_NBLK : 1;
    for (int r = 0; r < c->p.num_refs; r++)
        if (!subs[0]->ref_set[r]) return fail(c, JMME_ERR_STATE, "reference not set");
    // predictors: only the MB rows of this context's stripe are read (checked here, uploaded below) — rows
    // outside it may be left uninitialised by the caller
    const size_t per_mb = (size_t)npb * 2;
    if (pred)
        for (int r = 0; r < c->p.num_refs; r++) {
            const int16_t *q = pred + ((size_t)r * n_mb + (size_t)c->p.mb_row_begin * c->mb_w) * per_mb;
            const size_t n = (size_t)(c->p.mb_row_end - c->p.mb_row_begin) * c->mb_w * per_mb;
            for (size_t i = 0; i < n; i++)
                if (q[i] > c->max_pred || q[i] < -c->max_pred) return fail(c, JMME_ERR_PARAM, "pred out of range");
        }
    auto upload_pred = [&](jmme_ctx *s, cudaStream_t st) -> cudaError_t {      // the stripe rows of every reference
        for (int r = 0; pred && r < c->p.num_refs; r++) {
            const size_t off = ((size_t)r * n_mb + (size_t)s->p.mb_row_begin * s->mb_w) * per_mb;
            const size_t n = (size_t)(s->p.mb_row_end - s->p.mb_row_begin) * s->mb_w * per_mb;
            cudaError_t e = cudaMemcpyAsync(s->d_pred + off, pred + off, n * sizeof(int16_t), cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };

    // Single device, host buffers: the stripe is searched in parts on separate streams, so that the upload of
    // the later parts of the current picture and the download of the earlier parts of the MV field overlap
    // the kernels of the other parts (the reference planes are shared).
    if (ns == 1 && !c->profiling && c->tune.pipe_parts > 1 && c->p.pred_policy != JMME_PRED_MEDIAN && c->w16 == c->p.width && !(stride & 15) && !((uintptr_t)cur & 15) &&
        c->p.mb_row_end - c->p.mb_row_begin >= 4 * c->tune.pipe_parts) {
        jmme_ctx *s = c;
        CU(c, cudaSetDevice(s->device));
        CU(c, upload_pred(s, s->stream));
        CU(c, cudaEventRecord(s->ev_chunk[0], s->stream));     // first chunk of the planes (and predictors) ready after this
        const int np = s->tune.pipe_parts;
        int rc = JMME_OK, used = 0;
        // a failure in part k must not return while parts 0..k-1 still write into the caller's buffers
#define CUP(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) { rc = fail(c, JMME_ERR_CUDA, #call, e_); goto parts_done; }    \
    } while (0)
        for (int pt = 0; pt < np; pt++) {
            const int rb = part_begin(s, pt, np), re = part_begin(s, pt + 1, np);
            cudaStream_t st = s->part_stream[pt];
            const int y0 = std::min(16 * rb, s->p.height - 1), y1 = std::min(16 * re, s->p.height);
            used = pt + 1;
            CUP(upload_rows(s->d_raw + (size_t)y0 * s->p.width, cur + (size_t)y0 * stride, stride, s->p.width,
                            std::max(y1 - y0, 1), st));
            if (pt > 0) {
                // the reference chunks of this part, queued behind the part's current-picture rows (copy-engine order)
                for (int r = 0; r < s->p.num_refs; r++)
                    if (s->pend_next[r] && s->pend_next[r] <= pt) {
                        rc = queue_ref_chunks(s, r, s->pend_luma[r], s->pend_stride[r], pt, pt + 1, np);
                        if (rc != JMME_OK) goto parts_done;
                        s->pend_next[r] = pt + 1 < np ? pt + 1 : 0;
                    }
                CUP(cudaEventRecord(s->ev_chunk[pt], s->stream));
            }
            CUP(cudaStreamWaitEvent(st, s->ev_chunk[pt], 0));
            rc = enqueue_search(s, s->d_raw, s->p.width, s->d_pred, s->d_out, out_per_ref ? s->d_out_per_ref : nullptr,
                                st, rb, re);
            if (rc != JMME_OK) goto parts_done;
            const size_t off = (size_t)rb * s->mb_w, cnt = (size_t)(re - rb) * s->mb_w;
            CUP(cudaMemcpyAsync(out + off, s->d_out + off, cnt * sizeof(jmme_mbresult), cudaMemcpyDeviceToHost, st));
            if (out_per_ref)
                for (int r = 0; r < c->p.num_refs; r++)
                    CUP(cudaMemcpyAsync(out_per_ref + r * n_mb + off, s->d_out_per_ref + r * n_mb + off,
                                        cnt * sizeof(jmme_mbresult), cudaMemcpyDeviceToHost, st));
        }
#undef CUP
    parts_done:
        if (rc != JMME_OK) flush_pending_refs(s);              // (keeps the planes whole for the next call)
        {
            cudaError_t e = cudaStreamSynchronize(s->stream);   // the chunk uploads read the caller's reference buffers
            if (e != cudaSuccess && rc == JMME_OK) rc = fail(c, JMME_ERR_CUDA, "cudaStreamSynchronize(stream)", e);
        }
        for (int pt = 0; pt < used; pt++) {
            cudaError_t e = cudaStreamSynchronize(s->part_stream[pt]);
            if (e != cudaSuccess && rc == JMME_OK) rc = fail(c, JMME_ERR_CUDA, "cudaStreamSynchronize(part)", e);
        }
        return rc;
    }

    // enqueue on every device, then gather
    for (int g = 0; g < ns; g++) {
        jmme_ctx *s = subs[g];
        CU(c, cudaSetDevice(s->device));
        { int rcf = flush_pending_refs(s); if (rcf != JMME_OK) return fail(c, rcf, s->err); }
        {
            // only the current-picture rows of this stripe (row h-1 stands in for the replicated rows below it)
            const int s0 = std::min(16 * s->p.mb_row_begin, s->p.height - 1);
            const int s1 = std::min(16 * s->p.mb_row_end, s->p.height);
            // on its own stream: overlaps a plane kernel still running from jmme_set_reference (async_reference)
            CU(c, upload_rows(s->d_raw + (size_t)s0 * s->p.width, cur + (size_t)s0 * stride, stride, s->p.width,
                              std::max(s1 - s0, 1), s->copy_stream));
            CU(c, cudaEventRecord(s->ev_copy, s->copy_stream));
            CU(c, cudaStreamWaitEvent(s->stream, s->ev_copy, 0));
        }
        CU(c, upload_pred(s, s->stream));
        int rc = enqueue_search(s, s->d_raw, s->p.width, s->d_pred, s->d_out, out_per_ref ? s->d_out_per_ref : nullptr,
                                s->stream);
        if (rc != JMME_OK) return fail(c, rc, s->err);
        CU(c, cudaEventRecord(s->ev_done, s->stream));
    }
    jmme_ctx *s0 = subs[0];
    CU(c, cudaSetDevice(s0->device));
    for (int g = 1; g < ns; g++) {                // MV-field gather over NVLink: stripe g -> device 0
        jmme_ctx *s = subs[g];
        const size_t off = (size_t)s->p.mb_row_begin * s->mb_w, cnt = (size_t)(s->p.mb_row_end - s->p.mb_row_begin) * s->mb_w;
        CU(c, cudaStreamWaitEvent(s0->stream, s->ev_done, 0));
        CU(c, cudaMemcpyPeerAsync(s0->d_out + off, s0->device, s->d_out + off, s->device, cnt * sizeof(jmme_mbresult),
                                  s0->stream));
        if (out_per_ref)
            for (int r = 0; r < c->p.num_refs; r++)
                CU(c, cudaMemcpyPeerAsync(s0->d_out_per_ref + r * n_mb + off, s0->device,
                                          s->d_out_per_ref + r * n_mb + off, s->device, cnt * sizeof(jmme_mbresult),
                                          s0->stream));
    }
    const size_t off = (size_t)c->p.mb_row_begin * c->mb_w, cnt = (size_t)(c->p.mb_row_end - c->p.mb_row_begin) * c->mb_w;
    CU(c, cudaMemcpyAsync(out + off, s0->d_out + off, cnt * sizeof(jmme_mbresult), cudaMemcpyDeviceToHost, s0->stream));
    if (out_per_ref)
        for (int r = 0; r < c->p.num_refs; r++)
            CU(c, cudaMemcpyAsync(out_per_ref + r * n_mb + off, s0->d_out_per_ref + r * n_mb + off,
                                  cnt * sizeof(jmme_mbresult), cudaMemcpyDeviceToHost, s0->stream));
    CU(c, cudaStreamSynchronize(s0->stream));
    return JMME_OK;
}

}  // extern "C"
