// emul_core.cpp — host build of the per-thread device arithmetic (ya_vo_b200/csrc/fast_core.h,
// select_serial.h) plus a sequential model of the select kernel's rank-scatter partition.
// TEST INFRASTRUCTURE: lets `pytest -m "not gpu"` check the byte-SIMD segment test, the dp4a/dp2a
// Gaussian, the intrinsic-by-intrinsic Harris response and the std::sort replay against the
// oracle without a GPU.  Build with -ffp-contract=off.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../ya_vo_b200/csrc/fast_core.h"
#include "../../ya_vo_b200/csrc/select_serial.h"

static int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

// image padded by 4 pixels (reflect101) on every side, row pitch a multiple of 4, origin word-aligned
struct Padded {
    int H, W, pitch;
    std::vector<uint8_t> buf;
    Padded(const uint8_t *img, int H_, int W_) : H(H_), W(W_) {
        pitch = ((W + 8 + 4 + 3) / 4) * 4 + 4;
        buf.assign((size_t)(H + 8) * pitch, 0);
        for (int r = -4; r < H + 4; r++)
            for (int c = -4; c < W + 8 && c < pitch - 4; c++) {
                int rr = reflect101(r, H), cc = (c < W + 4) ? reflect101(c, W) : 0;
                buf[(size_t)(r + 4) * pitch + (c + 4)] = (c < W + 4) ? img[(size_t)rr * W + cc] : 0;
            }
    }
    // word holding pixel (r, x), x % 4 == 0
    const uint32_t *word(int r, int x) const {
        return reinterpret_cast<const uint32_t *>(buf.data() + (size_t)(r + 4) * pitch + (x + 4));
    }
};

extern "C" {

// corner mask (H x W bytes, 0/1) through the 4-pixel byte-SIMD path
void emul_fast_mask(const uint8_t *img, int H, int W, uint8_t *out) {
    Padded P(img, H, W);
    std::memset(out, 0, (size_t)H * W);
    for (int r = 4; r < H - 4; r++)
        for (int x = 0; x < W; x += 4) {
            bool pre;
            uint32_t nib = yavo_fast4(P.word(r - 3, x), P.word(r - 2, x), P.word(r - 1, x), P.word(r, x),
                                      P.word(r + 1, x), P.word(r + 2, x), P.word(r + 3, x), &pre);
            for (int b = 0; b < 4; b++) {
                int c = x + b;
                if (c >= 4 && c < W - 4 && ((nib >> b) & 1)) out[(size_t)r * W + c] = 1;
            }
        }
}

// the cheap necessary condition the detect kernel runs first: returns the number of pixels where the full test
// fires but the necessary condition does not (must be 0), and counts the quads each one lets through
int emul_fast_core_violations(const uint8_t *img, int H, int W, long long *core_quads, long long *full_quads) {
    Padded P(img, H, W);
    int bad = 0;
    *core_quads = *full_quads = 0;
    for (int r = 4; r < H - 4; r++)
        for (int x = 0; x < W; x += 4) {
            bool pre;
            const uint32_t nib = yavo_fast4(P.word(r - 3, x), P.word(r - 2, x), P.word(r - 1, x), P.word(r, x),
                                            P.word(r + 1, x), P.word(r + 2, x), P.word(r + 3, x), &pre);
            const uint32_t core = yavo_fast4_core(P.word(r, x), P.word(r + 1, x), P.word(r + 3, x));
            const uint32_t cn = (((core >> 7) & 0x01010101u) * 0x01020408u) >> 24;
            if (nib & ~cn) bad++;
            *core_quads += core != 0;
            *full_quads += nib != 0;
        }
    return bad;
}

float emul_harris(const uint8_t *img, int W, int row, int col) {
    int a, b, c;
    yavo_structure_tensor([&](int r, int cc) { return (int)img[(size_t)r * W + cc]; }, row, col, &a, &b, &c);
    return yavo_harris_from_tensor(a, b, c);
}

float emul_score_from_tensor(int a, int b, int c) { return yavo_harris_from_tensor(a, b, c); }

void emul_blur(const uint8_t *img, int H, int W, uint8_t *out) {
    Padded P(img, H, W);
    // horizontal pass for rows -4 .. H+4 (one spare row so every output row pair has 5 pairs)
    const int Wq = (W + 3) / 4 * 4;
    std::vector<uint32_t> h((size_t)(H + 10) * Wq, 0);
    for (int r = -4; r < H + 4; r++)
        for (int x = 0; x < Wq; x += 4) {
            const uint32_t *w = P.word(r, x);
            uint32_t o[4];
            yavo_blur_h4(w[-1], w[0], w[1], o);
            for (int b = 0; b < 4; b++) h[(size_t)(r + 4) * Wq + x + b] = o[b];
        }
    for (int r = 0; r < H; r += 2)
        for (int x = 0; x < W; x++) {
            uint32_t Pp[5];
            for (int k = 0; k < 5; k++) {
                uint32_t lo = h[(size_t)(r + 2 * k) * Wq + x];        // row r-4+2k
                uint32_t hi = h[(size_t)(r + 2 * k + 1) * Wq + x];    // row r-3+2k
                Pp[k] = lo | (hi << 16);
            }
            uint32_t o0, o1;
            yavo_blur_v2(Pp, &o0, &o1);
            out[(size_t)r * W + x] = (uint8_t)o0;
            if (r + 1 < H) out[(size_t)(r + 1) * W + x] = (uint8_t)o1;
        }
}

// ---- model of select_topk_kernel: same routing rules, same partition formulation ------------------
static const int SERIAL = YAVO_SORT_THRESHOLD;  // ranges of <= 16 elements: stable sort (the kernel's warp rank sort)

static int model_partition(yavo_ent *A, int f, int l) {
    yavo_median_to_first(A, f, l);
    const yavo_ent piv = A[f];
    const int n = l - f, cap = n / 2 + 1;
    std::vector<int> Lpos, Rpos;
    for (int i = 0; i < n - 1; i++) {
        if (!yavo_before(A[f + 1 + i], piv) && (int)Lpos.size() < cap) Lpos.push_back(f + 1 + i);
        if (!yavo_before(piv, A[l - 1 - i]) && (int)Rpos.size() < cap) Rpos.push_back(l - 1 - i);
    }
    const int nL = (int)Lpos.size(), nR = (int)Rpos.size();
    const int npairs = std::min(nL, nR);
    int m = 0;
    for (int i = 0; i < npairs; i++)
        if (Lpos[i] < Rpos[i]) {
            std::swap(A[Lpos[i]], A[Rpos[i]]);
            m++;
        }
    long long cut = 0x7fffffff;
    if (m < nL) cut = Lpos[m];
    if (m >= 1) cut = std::min<long long>(cut, Rpos[m - 1]);
    return (int)cut;
}

void emul_select(float *scores, int32_t *payload, int n, int K) {
    std::vector<yavo_ent> A(n);
    for (int i = 0; i < n; i++) A[i] = yavo_make_ent(scores[i], (uint32_t)payload[i]);
    struct R { int f, l, d; };
    std::vector<R> cur, nxt, serial;
    auto route = [&](std::vector<R> &q, int f, int l, int d) {
        if (f >= K || l - f <= 1) return;
        if (l - f <= SERIAL) serial.push_back({f, l, d});
        else if (d == 0) yavo_serial_heapsort(A.data(), f, l);
        else q.push_back({f, l, d});
    };
    if (n > 1) {
        int lg = 31 - __builtin_clz((unsigned)n);
        route(cur, 0, n, 2 * lg);
    }
    while (!cur.empty()) {
        nxt.clear();
        for (auto &r : cur) {
            int cut = model_partition(A.data(), r.f, r.l);
            route(nxt, cut, r.l, r.d - 1);
            route(nxt, r.f, cut, r.d - 1);
        }
        cur.swap(nxt);
    }
    for (auto &r : serial) std::stable_sort(A.begin() + r.f, A.begin() + r.l, [](yavo_ent a, yavo_ent b) { return yavo_before(a, b); });
    for (int i = 0; i < n; i++) {
        scores[i] = yavo_ent_score(A[i]);
        payload[i] = (int32_t)(uint32_t)A[i];
    }
}

// the serial pieces on their own: full introsort replay by one thread
void emul_serial_sort(float *scores, int32_t *payload, int n, int K, int depth_override) {
    std::vector<yavo_ent> A(n);
    for (int i = 0; i < n; i++) A[i] = yavo_make_ent(scores[i], (uint32_t)payload[i]);
    if (n > 1) {
        int lg = 31 - __builtin_clz((unsigned)n);
        yavo_serial_introsort(A.data(), 0, n, depth_override >= 0 ? depth_override : 2 * lg, K);
    }
    for (int i = 0; i < n; i++) {
        scores[i] = yavo_ent_score(A[i]);
        payload[i] = (int32_t)(uint32_t)A[i];
    }
}

}  // extern "C"
