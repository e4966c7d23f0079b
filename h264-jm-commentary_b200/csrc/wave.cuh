// wave.cuh — MV prediction pieces shared by predict.cu and the search kernels (in-frame median).
#pragma once
#include "jmme_dev.cuh"

struct Nb {          // one neighbour: vector, reference index (-1 = none), availability
    int x, y, ref, avail;
};

__host__ __device__ inline int med3(int a, int b, int c) { return max(min(a, b), min(max(a, b), c)); }

// 8.4.1.3: directional rules of 16x8 / 8x16, else 8.4.1.3.1 median rules.  C is already D when C is missing.
__host__ __device__ inline void mv_predict(int t, int part, int ref, Nb A, Nb B, Nb C, int &px, int &py)
{
    if (!A.avail || A.ref < 0) { A.x = A.y = 0; A.ref = -1; }
    if (!B.avail || B.ref < 0) { B.x = B.y = 0; B.ref = -1; }
    if (!C.avail || C.ref < 0) { C.x = C.y = 0; C.ref = -1; }
    if (t == 2 || t == 3) {                              // directional neighbour (by value: no pointers to locals)
        const Nb d = t == 2 ? (part == 0 ? B : A) : (part == 0 ? A : C);
        if (d.ref == ref) { px = d.x; py = d.y; return; }
    }
    if (!B.avail && !C.avail && A.avail) { B = A; C = A; }
    const int hit = (A.ref == ref) + (B.ref == ref) + (C.ref == ref);
    if (hit == 1) {
        const Nb &m = A.ref == ref ? A : (B.ref == ref ? B : C);
        px = m.x; py = m.y;
    } else {
        px = med3(A.x, B.x, C.x); py = med3(A.y, B.y, C.y);
    }
}

struct WaveNb { int16_t x, y; int16_t ref, avail; };
// where the neighbours A, B, C, D of each block come from: 0..9 = slot of the MB's outer neighbour cells
// (0..3 left column, 4..9 the row above from x-1), 10 = a partition of this MB decoded earlier (carries the
// 16x16 predictor), 11 = a partition of this MB decoded later (unavailable); tp = blocktype | part << 3
struct WaveTab { uint8_t src[JMME_NBLK][4]; uint8_t tp[JMME_NBLK]; };


// one outer neighbour cell (x, y) of an MB in field coordinates; y_top = first field row of the slice
__device__ __forceinline__ WaveNb wave_load_nb(const int16_t *mv4, const int8_t *ref4, int fw, int fh, int y_top, int x, int y)
{
    WaveNb v = {0, 0, -1, 0};                              // outside the picture or the slice: unavailable
    if (x >= 0 && x < fw && y >= y_top && y < fh) {
        const size_t o = (size_t)y * fw + x;
        const uint32_t w = *(const uint32_t *)(mv4 + 2 * o);
        v.ref = ref4[o]; v.avail = 1;
        if (v.ref >= 0) { v.x = (int16_t)(w & 0xFFFF); v.y = (int16_t)(w >> 16); }
    }
    return v;
}
// field position of outer-neighbour slot sl (0..3 left column, 4..9 the row above from x-1) of MB (mbx, mby)
__device__ __forceinline__ void wave_slot_xy(int mbx, int mby, int sl, int &x, int &y)
{
    x = sl < 4 ? 4 * mbx - 1 : 4 * mbx - 1 + (sl - 4);
    y = sl < 4 ? 4 * mby + sl : 4 * mby - 1;
}
// predictor of one block from the MB's 10 outer neighbour cells, its neighbour-source row and (p16, ref)
__device__ __forceinline__ void wave_predict_block(const WaveNb *nb10, const uint8_t *src4, int tp, int ref, int p16x,
                                                   int p16y, int &px, int &py)
{
    Nb nb[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int src = src4[q];
        const WaveNb v = nb10[min(src, 9)];
        if (src < 10) nb[q] = Nb{v.x, v.y, v.ref, v.avail};
        else if (src == 10) nb[q] = Nb{p16x, p16y, ref, 1};
        else nb[q] = Nb{0, 0, -1, 0};
    }
    if (!nb[2].avail) nb[2] = nb[3];
    mv_predict(tp & 7, tp >> 3, ref, nb[0], nb[1], nb[2], px, py);
}
