"""Independent numpy / pure-Python restatements used to cross-check the C oracle.
Written sample-by-sample from the H.264 text (8.4.2.2.1) and from the definitions in DESIGN.md §2,
deliberately NOT sharing structure with oracle/jmme_oracle.c."""
import numpy as np


def se_bits(v):
    """Signed Exp-Golomb code length: codeNum = 2|v| - (v>0); length = 2*floor(log2(codeNum+1))+1."""
    code = 2 * abs(v) - (1 if v > 0 else 0)
    return 2 * ((code + 1).bit_length() - 1) + 1


def ue_bits(v):
    return 2 * ((v + 1).bit_length() - 1) + 1


def spiral(R):
    """JM spiral_search_x/y, written as 'sort by ring then by the ring's emission order'."""
    pts = [(0, 0)]
    for l in range(1, R + 1):
        for i in range(-l + 1, l):
            pts += [(i, -l), (i, l)]
        for i in range(-l, l + 1):
            pts += [(-l, i), (l, i)]
    return pts


def spiral_index_closed_form(dx, dy):
    """SURVEY A.5 closed form."""
    if dx == 0 and dy == 0:
        return 0
    l = max(abs(dx), abs(dy))
    base = (2 * l - 1) ** 2
    if abs(dy) == l and abs(dx) < l:
        return base + 2 * (dx + l - 1) + (1 if dy > 0 else 0)
    return base + 2 * (2 * l - 1) + 2 * (dy + l) + (1 if dx > 0 else 0)


class Interp:
    """Per-sample luma interpolation on an infinitely edge-replicated picture."""

    def __init__(self, img):
        self.img = np.asarray(img, dtype=np.int64)
        self.H, self.W = self.img.shape

    def G(self, x, y):
        return int(self.img[min(max(y, 0), self.H - 1), min(max(x, 0), self.W - 1)])

    @staticmethod
    def tap(v):
        return v[0] - 5 * v[1] + 20 * v[2] + 20 * v[3] - 5 * v[4] + v[5]

    @staticmethod
    def clip(v):
        return min(max(v, 0), 255)

    def b1(self, x, y):   # between G(x,y) and G(x+1,y)
        return self.tap([self.G(x + i, y) for i in range(-2, 4)])

    def h1(self, x, y):   # between G(x,y) and G(x,y+1)
        return self.tap([self.G(x, y + i) for i in range(-2, 4)])

    def b(self, x, y):
        return self.clip((self.b1(x, y) + 16) >> 5)

    def h(self, x, y):
        return self.clip((self.h1(x, y) + 16) >> 5)

    def j(self, x, y):
        # horizontal 6-tap over vertical intermediates (the text allows either order)
        return self.clip((self.tap([self.h1(x + i, y) for i in range(-2, 4)]) + 512) >> 10)

    def sample(self, qx, qy):
        """Sample at quarter-pel position (qx, qy) relative to picture sample (0,0)."""
        x, y, fx, fy = qx >> 2, qy >> 2, qx & 3, qy & 3
        G, b, h, j = self.G, self.b, self.h, self.j
        avg = lambda p, q: (p + q + 1) >> 1                               # noqa: E731
        table = {
            (0, 0): lambda: G(x, y), (2, 0): lambda: b(x, y), (0, 2): lambda: h(x, y), (2, 2): lambda: j(x, y),
            (1, 0): lambda: avg(G(x, y), b(x, y)), (3, 0): lambda: avg(b(x, y), G(x + 1, y)),
            (0, 1): lambda: avg(G(x, y), h(x, y)), (0, 3): lambda: avg(h(x, y), G(x, y + 1)),
            (2, 1): lambda: avg(b(x, y), j(x, y)), (2, 3): lambda: avg(j(x, y), b(x, y + 1)),
            (1, 2): lambda: avg(h(x, y), j(x, y)), (3, 2): lambda: avg(j(x, y), h(x + 1, y)),
            (1, 1): lambda: avg(b(x, y), h(x, y)), (3, 1): lambda: avg(b(x, y), h(x + 1, y)),
            (1, 3): lambda: avg(h(x, y), b(x, y + 1)), (3, 3): lambda: avg(h(x + 1, y), b(x, y + 1)),
        }
        return table[(fx, fy)]()


H4 = np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, -1, 1], [1, -1, 1, -1]], dtype=np.int64)


def satd4x4(d, satd_round=0):
    d = np.asarray(d, dtype=np.int64).reshape(4, 4)
    s = int(np.abs(H4 @ d @ H4.T).sum())
    return (s + 1) >> 1 if satd_round else s >> 1


def weighted_cost(f, bits):
    return (f * bits) >> 16


def padded(img, pad):
    return np.pad(np.asarray(img), pad, mode="edge")


def brute_search(cur, refp, pad, bx, by, bw, bh, cx, cy, px, py, R, f, bonus=0, pretest=False):
    """Direct per-block search in spiral order with strict <.  Returns (mvx, mvy, cost, sad_surface)."""
    blk = cur[by:by + bh, bx:bx + bw].astype(np.int64)
    best = None
    cands = spiral(R)
    order = list(range(len(cands)))
    if pretest:
        order = [cands.index((-cx, -cy))] + order
    for pos in order:
        mx, my = cx + cands[pos][0], cy + cands[pos][1]
        r = refp[pad + by + my: pad + by + my + bh, pad + bx + mx: pad + bx + mx + bw].astype(np.int64)
        c = int(np.abs(blk - r).sum()) + weighted_cost(f, se_bits(4 * mx - px) + se_bits(4 * my - py))
        if mx == 0 and my == 0:
            c -= bonus
        if best is None or c < best[2]:
            best = (mx, my, c)
    return best


# ---- MV prediction (H.264 8.4.1.3) and the ME-only commit rule, restated independently --------------
def mv_predict(t, part, ref, A, B, Cn):
    """A, B, Cn = (mvx, mvy, ref, avail); Cn is D when C is unavailable."""
    def norm(n):
        return [n[0], n[1], n[2], n[3]] if (n[3] and n[2] >= 0) else [0, 0, -1, n[3]]
    A, B, Cn = norm(A), norm(B), norm(Cn)
    if t == 2 and part == 0 and B[2] == ref:
        return B[0], B[1]
    if t == 2 and part == 1 and A[2] == ref:
        return A[0], A[1]
    if t == 3 and part == 0 and A[2] == ref:
        return A[0], A[1]
    if t == 3 and part == 1 and Cn[2] == ref:
        return Cn[0], Cn[1]
    if not B[3] and not Cn[3] and A[3]:
        B, Cn = list(A), list(A)
    same = [n for n in (A, B, Cn) if n[2] == ref]
    if len(same) == 1:
        return same[0][0], same[0][1]
    return sorted([A[0], B[0], Cn[0]])[1], sorted([A[1], B[1], Cn[1]])[1]


def decode_rank(t, x, y):
    """Decoding order of the partition of blocktype t that contains pixel (x, y) of the MB."""
    from jmme import abi
    w, h = abi.BLC[t]
    if t <= 3:
        return (y // h) * (16 // w) + x // w
    return (2 * (y // 8) + x // 8) * 16 + ((y % 8) // h) * (8 // w) + (x % 8) // w


def predict_frame(mv4, ref4, mb_w, mb_h, num_refs, slice_rows=None):
    """slice_rows=None: predictors from a field of an earlier pass (jmme_predict_frame).
    slice_rows=k (0 = whole frame): the in-frame median policy (JMME_PRED_MEDIAN): rows above the slice are
    unavailable and the already-decoded partitions of the MB itself carry the MB's own 16x16 predictor."""
    from jmme import abi
    blocks = abi.block_table()
    pred = np.zeros((num_refs, mb_w * mb_h, 41, 2), np.int16)
    in_frame = slice_rows is not None
    k = (slice_rows or mb_h) if in_frame else mb_h
    for r in range(num_refs):
        for mb in range(mb_w * mb_h):
            mbx, mby = mb % mb_w, mb // mb_w
            top = (mby // k) * k if in_frame else 0
            for b, (t, x0, y0, w, h) in enumerate(blocks):
                def nb(px, py):                                    # neighbour at MB-relative pixel (px, py)
                    gx, gy = 16 * mbx + px, 16 * mby + py
                    if gx < 0 or gy < 0 or gx >= 16 * mb_w or gy >= 16 * mb_h:
                        return (0, 0, -1, 0)
                    nmx, nmy = gx // 16, gy // 16
                    if (nmy, nmx) > (mby, mbx) or nmy < top:
                        return (0, 0, -1, 0)                       # a later macroblock / another slice
                    if (nmy, nmx) == (mby, mbx):
                        if decode_rank(t, px, py) >= decode_rank(t, x0, y0):
                            return (0, 0, -1, 0)                   # a later partition of this macroblock
                        if in_frame:
                            return (int(pred[r, mb, 0, 0]), int(pred[r, mb, 0, 1]), r, 1)
                    return (int(mv4[gy // 4, gx // 4, 0]), int(mv4[gy // 4, gx // 4, 1]), int(ref4[gy // 4, gx // 4]), 1)
                A, B, Cc, D = nb(x0 - 1, y0), nb(x0, y0 - 1), nb(x0 + w, y0 - 1), nb(x0 - 1, y0 - 1)
                part = (y0 // 8) if t == 2 else ((x0 // 8) if t == 3 else 0)
                pred[r, mb, b] = mv_predict(t, part, r, A, B, Cc if Cc[3] else D)
    return pred


def commit_field(res, mb_w, mb_h, mask=0xFE):
    from jmme import abi
    blocks = abi.block_table()
    mv4 = np.zeros((4 * mb_h, 4 * mb_w, 2), np.int16)
    ref4 = np.zeros((4 * mb_h, 4 * mb_w), np.int8)
    mode = np.zeros((mb_w * mb_h, 5), np.uint8)
    INF = float("inf")
    for mb in range(mb_w * mb_h):
        cost = res[mb]["cost"].astype(np.int64)
        of_type = lambda t: [b for b, blk in enumerate(blocks) if blk[0] == t]              # noqa: E731
        J = [sum(cost[b] for b in of_type(t)) if (mask >> t) & 1 else INF for t in (1, 2, 3)]
        sub, j8 = [0] * 4, 0
        for q in range(4):
            cands = []
            for t in (4, 5, 6, 7):
                if (mask >> t) & 1:
                    inside = [b for b in of_type(t) if 2 * (blocks[b][2] // 8) + blocks[b][1] // 8 == q]
                    cands.append((sum(cost[b] for b in inside), t))
            if not cands:
                j8 = INF
                break
            c, sub[q] = min(cands)                                   # ties -> lower blocktype
            j8 += c
        J.append(j8)
        md = J.index(min(J))                                          # ties -> lower mode
        mode[mb] = [8 if md == 3 else md + 1] + ([*sub] if md == 3 else [0, 0, 0, 0])
        for cy in range(4):
            for cx in range(4):
                t = sub[2 * (cy // 2) + cx // 2] if md == 3 else md + 1
                b = next(b for b in of_type(t) if blocks[b][1] <= 4 * cx < blocks[b][1] + blocks[b][3]
                         and blocks[b][2] <= 4 * cy < blocks[b][2] + blocks[b][4])
                mbx, mby = mb % mb_w, mb // mb_w
                mv4[4 * mby + cy, 4 * mbx + cx] = res[mb]["mv"][b]
                ref4[4 * mby + cy, 4 * mbx + cx] = res[mb]["ref_idx"][b]
    return mv4, ref4, mode


# ---- round 2: cost domains, SSE / 8x8 Hadamard, chroma ME — restated from DESIGN.md §2 -------------------
H8 = np.kron(np.array([[1, 1], [1, -1]], dtype=np.int64), H4)          # any +-1 ordering: sum|.| is order-free


def satd8x8(d, satd_round=1):
    """JM HadamardSAD8x8: (sum |H8 D H8'| + 2) >> 2."""
    d = np.asarray(d, dtype=np.int64).reshape(8, 8)
    s = int(np.abs(H8 @ d @ H8.T).sum())
    return (s + 2) >> 2 if satd_round else s >> 2


def chroma_sample(img, x8, y8):
    """H.264 8.4.2.2.2 at eighth-pel position (x8, y8) of an infinitely edge-replicated chroma picture."""
    img = np.asarray(img)
    hh, ww = img.shape
    x, y, xf, yf = x8 >> 3, y8 >> 3, x8 & 7, y8 & 7
    g = lambda xx, yy: int(img[min(max(yy, 0), hh - 1), min(max(xx, 0), ww - 1)])      # noqa: E731
    return ((8 - xf) * (8 - yf) * g(x, y) + xf * (8 - yf) * g(x + 1, y) + (8 - xf) * yf * g(x, y + 1)
            + xf * yf * g(x + 1, y + 1) + 32) >> 6


def block_metric(a, b, metric, t8=False, satd_round=0):
    """Distortion of two equally sized blocks: 0 SAD, 1 SSE, 2 Hadamard (4x4 tiles, or 8x8 tiles when t8 and the
    block is at least 8x8; blocks thinner than 4 fall back to SAD)."""
    d = np.asarray(a, dtype=np.int64) - np.asarray(b, dtype=np.int64)
    bh, bw = d.shape
    if metric == 2 and (bw < 4 or bh < 4):
        metric = 0
    if metric == 0:
        return int(np.abs(d).sum())
    if metric == 1:
        return int((d * d).sum())
    n = 8 if (t8 and bw >= 8 and bh >= 8) else 4
    f = satd8x8 if n == 8 else satd4x4
    return sum(f(d[y:y + n, x:x + n], satd_round) for y in range(0, bh, n) for x in range(0, bw, n))


class StageSpec:
    """Per-stage metric and lambda factor of a context, cost-domain arithmetic."""

    def __init__(self, lam, domain=0, metrics=(0, 2, 2), t8=False, satd_round=0, chroma=False):
        self.domain, self.metrics, self.t8, self.satd_round, self.chroma = domain, metrics, t8, satd_round, chroma
        scale = 32.0 if domain else 65536.0
        self.lf = [int(scale * (lam * lam if m == 1 else lam) + 0.5) for m in metrics]

    def rate(self, st, bits):
        return self.lf[st] * bits if self.domain else (self.lf[st] * bits) >> 16

    def dist(self, d):
        return d << 5 if self.domain else d


def block_search(spec, cur, ref, bx, by, bw, bh, cx, cy, px, py, R, bonus16, pretest, subpel, cur_c=None, ref_c=None):
    """One block through all stages (integer spiral scan with strict <, then half- and quarter-pel), written
    directly from the definitions: per-sample interpolation (Interp / chroma_sample), no planes, no surfaces.
    bonus16: the block is the 16x16 block of reference 0 with !rdopt.  Returns (mvx, mvy, cost) in quarter-pel."""
    it = Interp(ref)
    blk = np.asarray(cur)[by:by + bh, bx:bx + bw]

    def luma_pred(qx, qy):
        return np.array([[it.sample(4 * (bx + x) + qx, 4 * (by + y) + qy) for x in range(bw)] for y in range(bh)])

    def cost_at(st, qx, qy):
        d = block_metric(blk, luma_pred(qx, qy), spec.metrics[st], spec.t8, spec.satd_round)
        if spec.chroma and st > 0:
            for k in range(2):
                cb = np.asarray(cur_c[k])[by // 2:(by + bh) // 2, bx // 2:(bx + bw) // 2]
                pr = np.array([[chroma_sample(ref_c[k], 8 * (bx // 2 + x) + qx, 8 * (by // 2 + y) + qy)
                                for x in range(bw // 2)] for y in range(bh // 2)])
                d += block_metric(cb, pr, spec.metrics[st], False, spec.satd_round)
        c = spec.dist(d) + spec.rate(st, se_bits(qx - px) + se_bits(qy - py))
        if bonus16 and qx == 0 and qy == 0:
            c -= spec.rate(st, 16)
        return c

    cands = spiral(R)
    order = list(range(len(cands)))
    if pretest:
        order = [cands.index((-cx, -cy))] + order
    best = None
    for pos in order:
        mx, my = cx + cands[pos][0], cy + cands[pos][1]
        c = cost_at(0, 4 * mx, 4 * my)
        if best is None or c < best[2]:
            best = (4 * mx, 4 * my, c)
    mvx, mvy, mn = best
    if not subpel:
        return mvx, mvy, mn
    sp = spiral(1)
    prev = spec.metrics[0]
    for step, st in ((2, 1), (1, 2)):
        restart = spec.chroma or spec.metrics[st] != prev
        if restart:
            mn = None
        ox, oy, bp = mvx, mvy, 0
        for pos in range(0 if restart else 1, 9):
            qx, qy = ox + step * sp[pos][0], oy + step * sp[pos][1]
            c = cost_at(st, qx, qy)
            if mn is None or c < mn:
                mn, bp = c, pos
        mvx, mvy = ox + step * sp[bp][0], oy + step * sp[bp][1]
        prev = spec.metrics[st]
    return mvx, mvy, mn


def bipred_block(spec, cur, refs0, ref1, bx, by, bw, bh, mv0, mv1, pr0, pr1, rng, iterations, pad):
    """Bi-predictive refinement of one block (include/jmme.h, jmme_search_frame_bipred), sample by sample through
    Interp: iteration i searches list i & 1 over the spiral of `rng` around its current vector, the other list
    fixed; prediction (a + b + 1) >> 1; rate of both vectors.  refs0: the list-0 picture.  Returns (mv0, mv1, cost)."""
    its = [Interp(refs0), Interp(ref1)]
    blk = np.asarray(cur, dtype=np.int64)[by:by + bh, bx:bx + bw]
    mv = [[int(v) for v in mv0], [int(v) for v in mv1]]
    pr = [[int(v) for v in pr0], [int(v) for v in pr1]]

    def block(k, qx, qy):
        return np.array([[its[k].sample(4 * (bx + x) + qx, 4 * (by + y) + qy) for x in range(bw)] for y in range(bh)], dtype=np.int64)

    cost = None
    for it in range(iterations):
        s, f = it & 1, 1 - (it & 1)
        fixed = block(f, mv[f][0], mv[f][1])
        fbits = se_bits(mv[f][0] - pr[f][0]) + se_bits(mv[f][1] - pr[f][1])
        best = None
        for dx, dy in spiral(rng):
            qx, qy = mv[s][0] + 4 * dx, mv[s][1] + 4 * dy
            if not (-(pad - 1) <= (qx >> 2) <= pad - 1 and -(pad - 1) <= (qy >> 2) <= pad - 1):
                continue
            e = blk - ((fixed + block(s, qx, qy) + 1) >> 1)
            d = int((e * e).sum()) if spec.metrics[0] == 1 else int(np.abs(e).sum())
            c = spec.dist(d) + spec.rate(0, fbits + se_bits(qx - pr[s][0]) + se_bits(qy - pr[s][1]))
            if best is None or c < best[0]:
                best = (c, qx, qy)
        cost, mv[s][0], mv[s][1] = best
    return tuple(mv[0]), tuple(mv[1]), cost
