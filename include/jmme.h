/*
 * jmme.h — C ABI of the JM (H.264 reference encoder, lencod) motion-estimation hot path.
 *
 * One header, two implementations:
 *   libjmme_cuda.so   the product: hand-written sm_100a kernels (h264-jm-commentary_b200/csrc/)
 *   libjmme_oracle.so the CPU oracle (oracle/), test infrastructure only
 *
 * Reference interface replaced: the mounted reference (/root/reference) holds only
 * README.md:1-4 and exposes no interface.  JM itself reaches motion
 * estimation through direct C calls; the entry points below carry the JM function each one
 * stands in for (names recalled from the public JM distribution, [MEM] in SURVEY.md — they
 * are NOT citations into /root/reference):
 *
 *   jmme_InitMotionSearchModule        Init_Motion_Search_Module / InitializeMotionSearch   (a1)
 *   jmme_lambda_factor                 LAMBDA_FACTOR(lambda_motion)                         (a2)
 *   jmme_set_reference                 UnifiedOneForthPix / getSubImagesLuma                (a12)
 *   jmme_getSubImagesLuma              getSubImagesLuma (stand-alone leaf)                  (a12)
 *   jmme_search_frame                  for every MB: PartitionMotionSearch -> BlockMotionSearch
 *                                        -> SetupFastFullPelSearch / FastFullPelBlockMotionSearch
 *                                        |  FullPelBlockMotionSearch -> SubPelBlockMotionSearch (a4-a10)
 *   jmme_SetupFastFullPelSearch        SetupFastFullPelSearch (+SetupLargerBlocks), one MB  (a6)
 *   jmme_FastFullPelBlockMotionSearch  FastFullPelBlockMotionSearch, one block              (a7)
 *   jmme_FullPelBlockMotionSearch      FullPelBlockMotionSearch, one block                  (a8)
 *   jmme_SubPelBlockMotionSearch       SubPelBlockMotionSearch, one block                   (a10)
 *   jmme_SATD                          SATD / HadamardSAD4x4, batched                       (a11)
 *   jmme_SetMotionVectorPredictor      SetMotionVectorPredictor / GetMotionVectorPredictorNormal (a3)
 *   jmme_predict_frame                 the same for all 41 blocks of every MB from a committed
 *                                      4x4-granular motion field (enc_picture->mv / ref_idx)
 *   jmme_commit_field                  ME-only stand-in for the mode decision that fills that field
 *
 * The arithmetic conventions ("the frozen spec") are in DESIGN.md §2.  C89-includable.
 */
#ifndef JMME_H
#define JMME_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JMME_ABI_VERSION     6
#define JMME_BLOCKS_PER_MB   41   /* 1 + 2 + 2 + 4 + 8 + 8 + 16 */
#define JMME_MAX_REFS        4
#define JMME_MAX_SEARCH_RANGE 64
#define JMME_MAX_GPUS        8
#define JMME_MAX_PRED_QPEL   2048 /* |predictor component| limit, quarter-pel: host entry points refuse larger */
                                  /* values (JMME_ERR_PARAM); the device-pointer entry points clamp to it       */

/* error codes: 0 = OK, negative otherwise; the library never exits or aborts */
#define JMME_OK               0
#define JMME_ERR_PARAM       (-1)
#define JMME_ERR_CUDA        (-2)
#define JMME_ERR_NOMEM       (-3)
#define JMME_ERR_UNSUPPORTED (-4)
#define JMME_ERR_STATE       (-5)   /* e.g. search before every reference was set */
#define JMME_ERR_NODEVICE    (-6)

/* search_mode */
#define JMME_SEARCH_FASTFULL 0      /* one window per MB/ref around the 16x16 predictor (a6,a7) */
#define JMME_SEARCH_FULL     1      /* one window per block around its own predictor (a8)      */

/* pred_policy: where the MV predictors (rate term + window centre) come from */
#define JMME_PRED_ZERO       0      /* (0,0) everywhere                                         */
#define JMME_PRED_PER_MB     1      /* caller passes one predictor per (ref, MB)                */
#define JMME_PRED_PER_BLOCK  2      /* caller passes 41 predictors per (ref, MB)                */
#define JMME_PRED_MEDIAN     3      /* closed loop inside the frame: H.264 8.4.1.3 median of the */
                                    /* MVs committed for the left / upper MBs of the same slice  */
                                    /* (params.slice_rows); DESIGN.md §2 "in-frame median"       */

/* blocktypes 1..7 = 16x16,16x8,8x16,8x8,8x4,4x8,4x4 (JM blc_size); bit t of blocktype_mask */
#define JMME_MASK_16x16      0x02
#define JMME_MASK_ALL        0xFE

typedef struct jmme_params {
    int32_t width, height;       /* luma size; padded up to x16 by right/bottom replication      */
    int32_t search_range;        /* R, 1..64; candidates = (2R+1)^2                              */
    int32_t num_refs;            /* 1..JMME_MAX_REFS                                             */
    int32_t blocktype_mask;      /* JMME_MASK_*                                                  */
    int32_t lambda_factor;       /* Q16 lambda_motion; 0 = derive from qp and rdopt              */
    int32_t qp;                  /* 0..51, used when lambda_factor == 0                          */
    int32_t rdopt;               /* 0: (0,0) pre-test + 16x16 (0,0) bonus + integer lambda       */
    int32_t use_hadamard;        /* sub-pel distortion: 1 = SATD, 0 = SAD                        */
    int32_t subpel;              /* 0 = integer search only, 1 = half- then quarter-pel          */
    int32_t search_mode;         /* JMME_SEARCH_*                                                */
    int32_t pred_policy;         /* JMME_PRED_*                                                  */
    int32_t satd_round;          /* 0: satd>>1 per 4x4 (Gen A), 1: (satd+1)>>1 (Gen B)           */
    int32_t cost_domain;         /* 0: JM <= 10: J = D + (lambda_factor*bits >> 16), lambda_factor Q16           */
                                 /* 1: JM >= 12 (JCOST_CALC_SCALEUP): J = (D << 5) + lambda_factor*bits,         */
                                 /*    lambda_factor = (int)(32*lambda + 0.5); every cost returned is in that    */
                                 /*    scaled-up domain                                           */
    int32_t mb_row_begin;        /* stripe of MB rows searched by this context: [begin, end)     */
    int32_t mb_row_end;          /* 0 = to the last row                                          */
    int32_t n_gpus;              /* 0/1 = one device; >1 = split the stripe over device_ids      */
    int32_t device_ids[JMME_MAX_GPUS];
    int32_t async_reference;     /* 1: jmme_set_reference returns once its copy and kernel are   */
                                 /* queued; the luma buffer must stay unchanged until the next  */
                                 /* jmme_search_frame / jmme_get_subimage returns (pinned memory) */
    int32_t slice_rows;          /* JMME_PRED_MEDIAN: MB rows per slice (neighbours outside the  */
                                 /* slice are unavailable); 0 = one slice = the frame.  Stripes  */
                                 /* (mb_row_begin/end, n_gpus) start and end on slice boundaries */
    /* ---- ABI 4: JM >= 12 distortion selection (MEDistortionFPel/HPel/QPel, Transform8x8Mode, ChromaMEEnable) --- */
    int32_t me_distortion;       /* 0: legacy — integer stage SAD, sub-pel stages SAD or (use_hadamard) SATD     */
                                 /* 1: the three fields below select the metric of each stage                    */
    int32_t me_distortion_fpel;  /* JMME_DIST_*: integer-pel stage (JMME_DIST_HADAMARD is refused there)         */
    int32_t me_distortion_hpel;  /* half-pel stage                                                               */
    int32_t me_distortion_qpel;  /* quarter-pel stage                                                            */
                                 /* a stage restarts at position 0 with the minimum reset when its metric        */
                                 /* differs from the previous stage's (or chroma_me is on); an SSE stage uses    */
                                 /* lambda^2 (JM: lambda_me = lambda_md for SSE)                                 */
    int32_t transform8x8;        /* 1: Hadamard distortion of blocktypes 1..4 (>= 8x8) is the sum of 8x8         */
                                 /* transforms, (sum|coef| + 2) >> 2 each (JM HadamardSAD8x8)                    */
    int32_t chroma_me;           /* 1: the sub-pel stages add the distortion of both chroma blocks (4:2:0,       */
                                 /* 1/8-pel bilinear samples, H.264 8.4.2.2.2); needs jmme_set_reference_chroma  */
                                 /* and jmme_set_current_chroma                                                  */
    /* ---- ABI 6: JM's rule for the search-window centre ----------------------------- */
    int32_t jm_center;           /* 0: centre = pred/4 clamped to +-R always (every window inside the default    */
                                 /*    replication border);  1: JM's BlockMotionSearch — clamped to +-R only     */
                                 /*    when !rdopt; with rdopt = 1 the centre follows the predictor, limited to  */
                                 /*    +-max_pred_qpel/4, and the planes get a border of max_pred_qpel/4 + R +   */
                                 /*    16 samples (jmme_pad) so that every window still lies inside them         */
    int32_t max_pred_qpel;       /* |predictor component| the host entry points accept, quarter-pel;             */
                                 /* 0 = JMME_MAX_PRED_QPEL (2048); 4..2048                                       */
} jmme_params;

/* distortion metrics (JM MEDistortion*: 0 SAD, 1 SSE, 2 Hadamard SAD) */
#define JMME_DIST_SAD        0
#define JMME_DIST_SSE        1
#define JMME_DIST_HADAMARD   2

/* One macroblock's result.  Block order: blocktype 1..7, raster order inside the MB
 * (index bases 0,1,3,5,9,17,25).  mv in quarter-pel units.  cost = distortion + MV rate
 * + reference rate of ref_idx.  Blocks whose blocktype is masked out: mv 0, cost INT32_MAX,
 * ref_idx -1. */
typedef struct jmme_mbresult {
    int16_t mv[JMME_BLOCKS_PER_MB][2];
    int32_t cost[JMME_BLOCKS_PER_MB];
    int8_t  ref_idx[JMME_BLOCKS_PER_MB];
    int8_t  reserved[3];
} jmme_mbresult;

typedef struct jmme_ctx jmme_ctx;

/* ---- life cycle ------------------------------------------------------------------------ */
void        jmme_default_params(jmme_params *p);
int         jmme_create(jmme_ctx **out, const jmme_params *p);
int         jmme_destroy(jmme_ctx *ctx);
const char *jmme_strerror(int code);
const char *jmme_last_error(const jmme_ctx *ctx);      /* detail of the last failure          */
const char *jmme_backend(void);                        /* "cuda-sm_100a" or "cpu-oracle"      */
int         jmme_abi_version(void);

/* geometry helpers (valid after create) */
int         jmme_mb_width(const jmme_ctx *ctx);
int         jmme_mb_height(const jmme_ctx *ctx);
int         jmme_pad(const jmme_ctx *ctx);             /* replication border of the ref planes */
int         jmme_lambda_factor_of(const jmme_ctx *ctx);
int         jmme_lambda_factor(int qp, int rdopt);     /* (int)(65536*lambda_motion + 0.5)     */

/* ---- frame-level path (host buffers; copies happen inside the call) -------------------- */
/* Upload reference `ref_idx`, replicate its borders and (when params.subpel) build the 16
 * quarter-pel planes. */
int jmme_set_reference(jmme_ctx *ctx, int ref_idx, const uint8_t *luma, int stride);

/* chroma_me: the chroma planes (4:2:0, width/2 x height/2 each) of reference `ref_idx`, and of the current
 * picture for the next jmme_search_frame.  Borders are replicated like the luma's. */
int jmme_set_reference_chroma(jmme_ctx *ctx, int ref_idx, const uint8_t *cb, const uint8_t *cr, int stride);
int jmme_set_current_chroma(jmme_ctx *ctx, const uint8_t *cb, const uint8_t *cr, int stride);

/* Search every MB of the context's stripe against every reference.
 *   pred         NULL for JMME_PRED_ZERO / JMME_PRED_MEDIAN, else int16 [num_refs][mb_count][nb][2] in
 *                quarter-pel units, nb = 1 (PER_MB) or 41 (PER_BLOCK); mb_count = whole frame
 *   out          [mb_w*mb_h] (whole-frame indexing; only the stripe's rows are written)
 *   out_per_ref  NULL or [num_refs][mb_w*mb_h]: per-reference winners, cost without reference
 *                rate (JM all_mv / motion_cost) */
int jmme_search_frame(jmme_ctx *ctx, const uint8_t *cur_luma, int stride,
                      const int16_t *pred, jmme_mbresult *out, jmme_mbresult *out_per_ref);

/* ---- (f2) bi-predictive refinement (JM BiPredMotionEstimation / BiPredMERefinements / BiPredMESearchRange) ------
 * List 0 = the context's references, list 1 = one more reference picture (jmme_set_reference_l1: uploaded and
 * interpolated like a list-0 reference).  jmme_search_frame_bipred refines, block by block, the vector pair
 * (mv0 towards reference l0.ref_idx of list 0, mv1 towards the list-1 picture) found by two uni-directional searches:
 *   iteration i = 0 .. iterations-1 searches list s = i & 1 with the other list's vector held fixed: candidates
 *   mv_s + 4 (dx, dy) over the spiral of `range` (position 0, the current pair, first; strict <), prediction =
 *   (P_fixed + P_candidate + 1) >> 1 per sample [STD 8.4.2.3 default weighted prediction] taken from the quarter-pel
 *   planes at each vector's own phase, cost = distortion (metric and cost domain of the integer stage) + the MV rate
 *   of BOTH vectors against their predictors; candidates whose integer displacement leaves +-(pad - 1) are skipped.
 * l0, l1: whole-frame jmme_search_frame outputs (host) of the list-0 context search and of a search against the
 * list-1 picture; pred0 / pred1: predictors of the two lists in the layout of the context's pred_policy (NULL = zero;
 * JMME_PRED_MEDIAN contexts take PER_BLOCK arrays here).  range 1..15, iterations 1..8.  Blocks whose blocktype is
 * masked out: vectors 0, cost INT32_MAX, ref0 -1.  Only the stripe's rows are written.  A context that searches a
 * stripe holds the list-0 planes for the rows its own search can reach (+-(2 R + 4) around the stripe): the product
 * library answers JMME_ERR_UNSUPPORTED when range x iterations could leave them. */
typedef struct jmme_bipred {
    int16_t mv0[JMME_BLOCKS_PER_MB][2];
    int16_t mv1[JMME_BLOCKS_PER_MB][2];
    int32_t cost[JMME_BLOCKS_PER_MB];
    int8_t  ref0[JMME_BLOCKS_PER_MB];
    int8_t  reserved[3];
} jmme_bipred;
int jmme_set_reference_l1(jmme_ctx *ctx, const uint8_t *luma, int stride);
int jmme_search_frame_bipred(jmme_ctx *ctx, const uint8_t *cur_luma, int stride, const jmme_mbresult *l0,
                             const jmme_mbresult *l1, const int16_t *pred0, const int16_t *pred1, int range,
                             int iterations, jmme_bipred *out);

/* The predictors the last jmme_search_frame of a JMME_PRED_MEDIAN context used:
 * int16 [num_refs][mb_w*mb_h][41][2] (only the stripe's rows are meaningful). */
int jmme_get_predictors(jmme_ctx *ctx, int16_t *pred);

/* Copy one quarter-pel plane (xfrac,yfrac in 0..3) of reference ref_idx back to the host,
 * padded size (W16+2*pad) x (H16+2*pad), for parity checks of (a12). */
int jmme_get_subimage(jmme_ctx *ctx, int ref_idx, int xfrac, int yfrac,
                      uint8_t *dst, int dst_stride);

/* ---- device-resident variants (product library only; pointers are CUDA device pointers,
 *      stream is a cudaStream_t passed as void*; asynchronous on that stream) -------------- */
int jmme_set_reference_dev(jmme_ctx *ctx, int ref_idx, const void *d_luma, int stride, void *stream);
int jmme_search_frame_dev(jmme_ctx *ctx, const void *d_cur_luma, int stride, const void *d_pred,
                          void *d_out, void *d_out_per_ref, void *stream);
int jmme_set_reference_chroma_dev(jmme_ctx *ctx, int ref_idx, const void *d_cb, const void *d_cr, int stride, void *stream);
int jmme_set_current_chroma_dev(jmme_ctx *ctx, const void *d_cb, const void *d_cr, int stride, void *stream);
/* Multi-GPU gather without a collective library: copy this context's stripe of the MV field from
 * d_field_local (whole-frame indexed, as written by jmme_search_frame_dev) into the same offsets of
 * n_peers peer buffers — device pointers mapped into this process (CUDA IPC / symmetric memory); the
 * stores travel over NVLink.  An entry equal to d_field_local is skipped.  Asynchronous on `stream`; the
 * caller provides the cross-rank barrier before the field is read. */
int jmme_push_stripe_dev(jmme_ctx *ctx, const void *d_field_local, void *const *d_field_peers, int n_peers,
                         void *stream);
/* The same gather fused into the search: after this call every record jmme_search_frame_dev writes into its
 * output buffer is also stored, by the kernel that produces it, into the same offsets of the peer buffers
 * (d_field_peers as in jmme_push_stripe_dev; entries that are NULL or equal to the output buffer of the search
 * are skipped; n_peers = 0 turns the fusion off).  No push kernel is needed then.  The caller provides the
 * cross-rank barriers: before the search (the peers have finished reading the previous field) and after it
 * (every stripe has landed everywhere). */
int jmme_set_peer_fields_dev(jmme_ctx *ctx, void *const *d_field_peers, int n_peers);
/* The fused gather through an NVLS multicast mapping: d_field_multicast is the multicast address of the symmetric
 * field the searches of every rank write into (e.g. torch.distributed._symmetric_memory: handle.multicast_ptr).  Every
 * record is then stored ONCE with multimem.st and the NVSwitch delivers it to the field of every rank (its own
 * included), instead of one peer store per rank; takes precedence over jmme_set_peer_fields_dev.  NULL turns it off.
 * The caller provides the same two cross-rank barriers. */
int jmme_set_multicast_field_dev(jmme_ctx *ctx, void *d_field_multicast);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
int64_t jmme_launch_count(const jmme_ctx *ctx);

/* Launch tuning of the product library.  Explicit per-context state (no environment variables): every field
 * 0 = the library's measured default.  Results never depend on it — only which instantiation
 * of the kernels runs and how the work is cut; the parity tests sweep it.  jmme_set_tuning applies to later
 * searches of the context (and to every sub-context of an n_gpus parent). */
typedef struct jmme_tuning {
    int32_t variant;         /* integer-search kernel: 10*K + launch shape (me_int.cu / me_int_tb.cu); 0 = by range */
    int32_t group;           /* zero-predictor items of 1, 2 or 4 adjacent MBs sharing one window; 0 = by mode       */
    int32_t cluster;         /* largest thread-block cluster of a wavefront step: 1, 2, 4; 0 = default (4)           */
    int32_t table_rate;      /* 1: per-block rate always from the table (no linear-rate form for integer lambda)    */
    int32_t wave_step;       /* 1: in-frame median predictors always by the separate wave_step kernel                */
    int32_t no_pdl;          /* 1: no programmatic dependent launch between the kernels of a wavefront step         */
    int32_t pipe_parts;      /* host path: the stripe is searched in this many parts on separate streams (1..4)      */
    int32_t balance;         /* zero-predictor search, R = 32: 1 = balanced task ranges (every CTA an equal range of */
                             /* the stripe's tasks); 0 / 2 = whole MB items per CTA (the measured default)           */
    int32_t even_parts;      /* host path: 1 = parts of equal size (default: a small first and a smaller last part)  */
    int32_t early_subpel;    /* sub-pel kernel as a programmatic dependent that waits per MB for the integer result  */
                             /* (runs beside the last round of the search kernel): 0 = when the search deals whole   */
                             /* items, 2 = never                                                                     */
    int32_t no_pair_tail;    /* 1: items of 4 MBs to the last row (default: whole rounds of them, then rows of pairs) */
    int32_t reserved[5];
} jmme_tuning;
int jmme_set_tuning(jmme_ctx *ctx, const jmme_tuning *t);
int jmme_get_tuning(const jmme_ctx *ctx, jmme_tuning *t);      /* the values in effect (defaults resolved)         */
/* Name and template arguments of the integer-search kernel the last search of this context launched, e.g.
 * "me_int_tb_kernel<K=6,NW=4,MINB=3,PER_BLOCK=0,RS_CT=126,KEYG=0,KRTAB=1,NMB=4,CL=1,WP=0,LIN=0,BAL=0>" ("" before the
 * first search; the oracle returns "cpu-oracle").  The parity tests assert it so that a test of a BASELINE
 * config provably ran the kernel the bench times. */
const char *jmme_last_kernel(const jmme_ctx *ctx);
/* Per-kernel device times.  jmme_set_profiling(ctx,1) makes every later set_reference / search
 * bracket its kernels with CUDA events on the launching stream; jmme_get_kernel_times waits for
 * the last bracket and returns milliseconds of the most recent
 *   ms[0] quarter-pel plane kernel (a12)   ms[1] integer search kernel (a6,a7)
 *   ms[2] sub-pel kernel (a10,a11)         ms[3] reference-selection kernel
 * (0 for a kernel that did not run).  Single-device contexts only. */
int jmme_set_profiling(jmme_ctx *ctx, int enable);
int jmme_get_kernel_times(jmme_ctx *ctx, float ms[4]);

/* ---- JM-named leaf entry points on plain arrays ----------------------------------------- */
/* (a1) tables.  mvbits has 2*max_mvd+1 entries, index (v + max_mvd); refbits n_refbits;
 * spiral_x/y (2R+1)^2 entries. Any pointer may be NULL. */
int jmme_InitMotionSearchModule(int search_range, int max_mvd, int32_t *mvbits,
                                int n_refbits, int32_t *refbits,
                                int16_t *spiral_x, int16_t *spiral_y);

/* (a3) H.264 8.4.1.3 median / directional MV prediction of one partition from its neighbours.
 * blocktype 1..7; part: 16x8 0 = upper, 1 = lower; 8x16 0 = left, 1 = right; ignored otherwise.
 * Neighbour C must already be D when C is not available.  refN < 0 = intra / other list. */
int jmme_SetMotionVectorPredictor(int blocktype, int part, int ref_idx,
                                  const int16_t mvA[2], int refA, int availA,
                                  const int16_t mvB[2], int refB, int availB,
                                  const int16_t mvC[2], int refC, int availC, int16_t pred[2]);

/* ME-only mode decision: per MB the cheapest of 16x16, 16x8, 8x16 and P8x8 (each 8x8
 * the cheapest of 8x8, 8x4, 4x8, 4x4) by summed block cost, lower mode on ties; writes the chosen
 * vectors into a 4x4-granular field: mv4 [4*mb_h][4*mb_w][2], ref4 [4*mb_h][4*mb_w],
 * mode [mb_h*mb_w][5] = {MB mode 1,2,3 or 8, sub-type of the four 8x8s (0 when unused)}.
 * res: whole-frame jmme_search_frame output (host). */
int jmme_commit_field(jmme_ctx *ctx, const jmme_mbresult *res, int16_t *mv4, int8_t *ref4, uint8_t *mode);

/* MV predictors of all 41 blocks of every MB and reference from a committed field: neighbours
 * A (left), B (up), C (up-right, else D up-left) with the standard's availability (picture
 * border, MB raster order, partition order inside the MB), then jmme_SetMotionVectorPredictor.
 * pred: int16 [num_refs][mb_h*mb_w][41][2], the layout JMME_PRED_PER_BLOCK takes. */
int jmme_predict_frame(jmme_ctx *ctx, const int16_t *mv4, const int8_t *ref4, int16_t *pred);

/* (a12) 16 planes from one luma image; out[(yfrac*4+xfrac)] planes are contiguous, each
 * (width+2*pad) x (height+2*pad) bytes; width/height must be multiples of 16. */
int jmme_getSubImagesLuma(const uint8_t *luma, int width, int height, int stride, int pad,
                          uint8_t *out_planes);

/* (a11) n 4x4 difference blocks (raster, int16) -> n SATD values */
int jmme_SATD(const int16_t *diff4x4, int n, int satd_round, int32_t *out);

/* (f2) JM HadamardSAD8x8: n 8x8 difference blocks (raster, int16) -> (sum|H8 D H8'| + 2) >> 2 each
 * (satd_round = 0: >> 2 without the rounding offset) */
int jmme_HadamardSAD8x8(const int16_t *diff8x8, int n, int satd_round, int32_t *out);

/* (f3) JM getSubImagesChroma for 4:2:0: the 64 eighth-pel planes of one chroma component, H.264 8.4.2.2.2:
 * ((8-xF)(8-yF)A + xF(8-yF)B + (8-xF)yF C + xF yF D + 32) >> 6 with edge replication.  out[(yF*8+xF)] planes
 * are contiguous, each (width+2*pad) x (height+2*pad) bytes (width, height: the chroma size). */
int jmme_getSubImagesChroma(const uint8_t *chroma, int width, int height, int stride, int pad, uint8_t *out_planes);

/* (a6) SAD surfaces of one MB: out[41][(2R+1)^2] in spiral order around centre (cx,cy)
 * (integer pel), including the 16x16 (0,0) bonus when bonus != 0 (subtracted at MV (0,0)).
 * ref_padded points at sample (0,0) of a plane with `pad` replicated pixels around it. */
int jmme_SetupFastFullPelSearch(const uint8_t *cur_mb16x16, int cur_stride,
                                const uint8_t *ref_padded, int ref_stride,
                                int mb_x, int mb_y, int cx, int cy, int search_range,
                                int bonus, int32_t *out_blocksad);

/* (a7) argmin over one block's SAD surface with the MV rate term.  pretest00: test MV (0,0)
 * first (JM !rdopt).  Returns the winning integer MV and cost. */
int jmme_FastFullPelBlockMotionSearch(const int32_t *blocksad, int search_range,
                                      int cx, int cy, int pred_x, int pred_y,
                                      int lambda_factor, int pretest00,
                                      int16_t *mv_x, int16_t *mv_y, int32_t *min_mcost);

/* (a8) per-block full search.  block_x/block_y: luma position of the block, bw x bh size. */
int jmme_FullPelBlockMotionSearch(const uint8_t *cur, int cur_stride,
                                  const uint8_t *ref_padded, int ref_stride,
                                  int block_x, int block_y, int bw, int bh,
                                  int pred_x, int pred_y, int search_range,
                                  int lambda_factor, int bonus,
                                  int16_t *mv_x, int16_t *mv_y, int32_t *min_mcost);

/* (a10) sub-pel refinement of one block.  planes: 16 padded planes as produced by
 * jmme_getSubImagesLuma.  In: integer-pel MV*4 and its cost; out: quarter-pel MV and cost. */
int jmme_SubPelBlockMotionSearch(const uint8_t *cur, int cur_stride,
                                 const uint8_t *planes, int width, int height, int pad,
                                 int block_x, int block_y, int bw, int bh,
                                 int pred_x, int pred_y, int lambda_factor,
                                 int use_hadamard, int satd_round, int bonus,
                                 int16_t *mv_x, int16_t *mv_y, int32_t *min_mcost);

#ifdef __cplusplus
}
#endif
#endif /* JMME_H */
