/* jmme_demo.c — the C ABI of include/jmme.h from plain C89 host code (the way a JM build would call it).
 *
 *   cc -std=c89 -pedantic -Iinclude examples/jmme_demo.c -Lh264-jm-commentary_b200/csrc -ljmme_cuda \
 *      -Wl,-rpath,$PWD/h264-jm-commentary_b200/csrc -o jmme_demo && ./jmme_demo
 *
 * Builds a synthetic reference, a current picture that is the reference displaced by (+3, -2) integer samples,
 * searches it with zero predictors and with the in-frame median policy, and prints a checksum of the motion
 * field.  Linked against the CPU oracle instead (tests/test_c_demo.py) it must print the same lines. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jmme.h"

#define W 96
#define H 64
#define R 8

static unsigned long lcg(unsigned long *s) { *s = (*s * 1103515245UL + 12345UL) & 0x7fffffffUL; return *s >> 8; }

static unsigned long checksum(const jmme_mbresult *r, int n)
{
    unsigned long h = 2166136261UL;
    int i, b;
    for (i = 0; i < n; i++)
        for (b = 0; b < JMME_BLOCKS_PER_MB; b++) {
            h = (h ^ (unsigned long)(r[i].mv[b][0] & 0xffff)) * 16777619UL & 0xffffffffUL;
            h = (h ^ (unsigned long)(r[i].mv[b][1] & 0xffff)) * 16777619UL & 0xffffffffUL;
            h = (h ^ (unsigned long)(r[i].cost[b] & 0xffffffffL)) * 16777619UL & 0xffffffffUL;
            h = (h ^ (unsigned long)(r[i].ref_idx[b] & 0xff)) * 16777619UL & 0xffffffffUL;
        }
    return h;
}

static int run(int policy, const unsigned char *ref, const unsigned char *cur)
{
    jmme_params p;
    jmme_ctx *ctx = NULL;
    jmme_mbresult *out;
    int rc, n, i, hits = 0;
    jmme_default_params(&p);
    p.width = W; p.height = H; p.search_range = R; p.num_refs = 1; p.qp = 28; p.subpel = 1;
    p.pred_policy = policy; p.slice_rows = 2;
    rc = jmme_create(&ctx, &p);
    if (rc != JMME_OK) { printf("jmme_create: %s\n", jmme_strerror(rc)); return 1; }
    n = jmme_mb_width(ctx) * jmme_mb_height(ctx);
    out = (jmme_mbresult *)calloc((size_t)n, sizeof *out);
    if (!out) return 1;
    rc = jmme_set_reference(ctx, 0, ref, W);
    if (rc == JMME_OK) rc = jmme_search_frame(ctx, cur, W, NULL, out, NULL);
    if (rc != JMME_OK) { printf("search: %s (%s)\n", jmme_strerror(rc), jmme_last_error(ctx)); return 1; }
    for (i = 0; i < n; i++) hits += out[i].mv[0][0] == 12 && out[i].mv[0][1] == -8;      /* quarter-pel units */
    printf("policy %d: %d MBs, 16x16 vector (+3,-2) found in %d, field checksum %08lx\n", policy, n, hits, checksum(out, n));
    free(out);
    jmme_destroy(ctx);
    return 0;
}

int main(void)
{
    static unsigned char ref[H][W], cur[H][W];
    unsigned long s = 7;
    int x, y;
    for (y = 0; y < H; y++)
        for (x = 0; x < W; x++) ref[y][x] = (unsigned char)((lcg(&s) & 63) + 2 * ((x / 8 + y / 8) & 15) + 64);
    for (y = 0; y < H; y++)
        for (x = 0; x < W; x++) {
            int sx = x + 3, sy = y - 2;
            sx = sx < 0 ? 0 : (sx >= W ? W - 1 : sx);
            sy = sy < 0 ? 0 : (sy >= H ? H - 1 : sy);
            cur[y][x] = ref[sy][sx];
        }
    printf("backend %s, ABI %d\n", jmme_backend(), jmme_abi_version());
    if (run(JMME_PRED_ZERO, &ref[0][0], &cur[0][0])) return 1;
    if (run(JMME_PRED_MEDIAN, &ref[0][0], &cur[0][0])) return 1;
    return 0;
}
