// yavo_oracle.cpp — CPU ORACLE (test infrastructure only; see yavo_oracle.h).
//
// A restatement of the reference's hot path, written for clarity and for
// bit-identical results, not copied from it.  Every function cites the
// reference lines it follows (paths relative to the reference checkout).
// Build: g++ -O2 -ffp-contract=off (no -ffast-math) so that no float
// expression is contracted to FMA — the reference's CMake sets no flags.
#include "yavo_oracle.h"

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <set>
#include <thread>
#include <utility>
#include <vector>

namespace {

// Net effect of src/FastDetector.cc:50-112 for radius 3 (the comparators at :19-45
// order and de-duplicate by the second coordinate only).  (dx, dy): dx is added to
// the first argument (row), dy to the second (col).  Checked against
// yavo_oracle_ring_literal and against tests/testBresenham.png.
const int kRing[16][2] = {{0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0},  {3, 1},  {2, 2},   {1, 3},
                          {0, 3},  {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3}};

const int kIntensityThreshold = 40;  // include/FastDetector.hpp:35 (ctor argument ignored)
const int kRunLength = 12;           // src/FastDetector.cc:147 (literal)

inline bool in_between(int c, int p) {  // src/FastDetector.cc:155-161
    return (c > p - kIntensityThreshold) && (c < p + kIntensityThreshold);
}

// OpenCV modules/core/src/lapack.cpp hypot<_Tp> (float instantiation)
inline float cv_hypot(float a, float b) {
    a = std::fabs(a);
    b = std::fabs(b);
    if (a > b) {
        b /= a;
        return a * std::sqrt(1 + b * b);
    }
    if (b > 0) {
        a /= b;
        return b * std::sqrt(1 + a * a);
    }
    return 0;
}

// OpenCV JacobiImpl_<float> specialised to n = 2 (one rotation, then the
// descending sort).  cv::eigen dispatches here when OpenCV is built without Eigen.
inline void eigen2x2(float a, float b, float c, float *l1, float *l2) {
    float W0 = a, W1 = c;
    const float eps = 1.1920928955078125e-07f;  // FLT_EPSILON
    float p = b;
    if (!(std::fabs(p) <= eps)) {
        float y = (float)((W1 - W0) * 0.5);
        float t = std::fabs(y) + cv_hypot(p, y);
        t = (p / t) * p;
        if (y < 0) t = -t;
        W0 -= t;
        W1 += t;
    }
    if (W0 < W1) std::swap(W0, W1);
    *l1 = W0;
    *l2 = W1;
}

// src/FastDetector.cc:270: float*float product; std::pow(float,int) promotes to double;
// 0.04 is a double literal; the double result is narrowed to float on return.
inline float harris_from_eigen(float l1, float l2) {
    float prod = l1 * l2;
    float sum = l2 + l1;
    double sq = (double)sum * (double)sum;  // == std::pow((double)sum, 2): exact in double
    double r = (double)prod - 0.04 * sq;
    return (float)r;
}

// Sobel responses at (r,c) as src/FastDetector.cc:164-200 computes them for interior
// pixels (3x3 correlation, no kernel flip).  Exact integers.
inline void sobel_at(const uint8_t *img, int W, int r, int c, int *gx, int *gy) {
    const uint8_t *p0 = img + (size_t)(r - 1) * W + c;
    const uint8_t *p1 = p0 + W;
    const uint8_t *p2 = p1 + W;
    *gx = (p0[1] - p0[-1]) + 2 * (p1[1] - p1[-1]) + (p2[1] - p2[-1]);
    *gy = (p2[-1] + 2 * p2[0] + p2[1]) - (p0[-1] + 2 * p0[0] + p0[1]);
}

// src/FastDetector.cc:244-273 for a pixel whose 5x5 neighbourhood is inside the image
// (always true for FAST candidates: x in [4,H-5], y in [4,W-5]).
inline float harris_interior(const uint8_t *img, int W, int x, int y) {
    // float accumulation of exact integers < 2^24 is exact, so integer sums give the same M
    int a = 0, b = 0, c = 0;
    for (int i = x - 1; i <= x + 1; i++)
        for (int j = y - 1; j <= y + 1; j++) {
            int gx, gy;
            sobel_at(img, W, i, j, &gx, &gy);
            a += gx * gx;
            b += gx * gy;
            c += gy * gy;
        }
    float l1, l2;
    eigen2x2((float)a, (float)b, (float)c, &l1, &l2);
    return harris_from_eigen(l1, l2);
}

inline bool is_corner(const uint8_t *img, int W, int i, int j) {  // src/FastDetector.cc:300-320
    const uint8_t *p = img + (size_t)i * W + j;
    int c = *p;
    auto ringv = [&](int k) { return (int)p[kRing[k][0] * W + kRing[k][1]]; };
    if (in_between(c, ringv(0)) || in_between(c, ringv(7))) return false;   // :315
    if (in_between(c, ringv(4)) && in_between(c, ringv(12))) return false;  // :317
    int run = 0;                                                            // :135-153
    for (int k = 0; k < 16; k++) {
        if (in_between(c, ringv(k)))
            run = 0;
        else
            run++;
        if (run >= kRunLength) return true;
    }
    return false;
}

struct Cand {
    int x, y;
    float score;
};

void collect_candidates(const uint8_t *img, int H, int W, std::vector<Cand> &out) {
    for (int i = 4; i < H - 4; i++)
        for (int j = 4; j < W - 4; j++)
            if (is_corner(img, W, i, j)) out.push_back({i, j, harris_interior(img, W, i, j)});
}

// OpenCV 4.x fixed-point GaussianBlur for CV_8U, ksize 9, sigma 2.5:
// getGaussianKernelBitExact -> 8 fractional bits, sum 256.
const int kGauss[9] = {12, 22, 31, 41, 44, 41, 31, 22, 12};

inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0)
            p = -p;
        else
            p = 2 * (n - 1) - p;
    }
    return p;
}

void gaussian_blur(const uint8_t *img, int H, int W, uint8_t *out) {
    std::vector<uint16_t> h((size_t)H * W);
    for (int r = 0; r < H; r++) {
        const uint8_t *row = img + (size_t)r * W;
        for (int c = 0; c < W; c++) {
            unsigned s = 0;
            if (c >= 4 && c + 4 < W) {
                for (int k = 0; k < 9; k++) s += kGauss[k] * row[c + k - 4];
            } else {
                for (int k = 0; k < 9; k++) s += kGauss[k] * row[reflect101(c + k - 4, W)];
            }
            h[(size_t)r * W + c] = (uint16_t)s;  // <= 255*256 fits u16
        }
    }
    for (int r = 0; r < H; r++) {
        const uint16_t *rows[9];
        for (int k = 0; k < 9; k++) rows[k] = h.data() + (size_t)reflect101(r + k - 4, H) * W;
        for (int c = 0; c < W; c++) {
            unsigned v = 0;
            for (int k = 0; k < 9; k++) v += kGauss[k] * rows[k][c];
            out[(size_t)r * W + c] = (uint8_t)((v + 32768u) >> 16);
        }
    }
}

inline bool check_boundary(int x, int y, int width, int height) {  // src/BriefDescriptor.cc:128-136
    if (x - 8 < 0 || x + 8 > width) return false;
    if (y - 8 < 0 || y + 8 > height) return false;
    return true;
}

void brief(const uint8_t *S, int H, int W, const int32_t *off, const int32_t *rows,
           const int32_t *cols, int n, uint8_t *desc, uint8_t *valid, int *n_oob) {
    const long long total = (long long)H * W;
    int oob = 0;
    for (int i = 0; i < n; i++) {
        uint8_t *d = desc + (size_t)i * 32;
        std::memset(d, 0, 32);
        int row = rows[i], col = cols[i];
        bool ok = check_boundary(col, row, W, H);  // :97 passes (y, x)
        valid[i] = ok ? 1 : 0;
        if (!ok) continue;
        bool touched = false;
        for (int j = 0; j < 256; j++) {
            // Image::getPixelVal(i,j) = data[i*cols + j] (src/Image.cc:15-17): linear, unchecked
            long long ia = (long long)(row + off[4 * j + 0]) * W + (col + off[4 * j + 1]);
            long long ib = (long long)(row + off[4 * j + 2]) * W + (col + off[4 * j + 3]);
            int va = 0, vb = 0;
            if (ia < total) va = S[ia]; else touched = true;  // reference UB: defined as 0
            if (ib < total) vb = S[ib]; else touched = true;
            if (va > vb) d[j / 8] |= (uint8_t)(1u << (j % 8));  // :108-117
        }
        if (touched) oob++;
    }
    if (n_oob) *n_oob = oob;
}

inline int popcount_serial(uint8_t v) {  // src/BriefDescriptor.cc:151-160
    int count = 0;
    while (v != 0) {
        if (v & 1) count++;
        v = (uint8_t)(v >> 1);
    }
    return count;
}

inline int hamming_fast(const uint8_t *a, const uint8_t *b) {
    uint64_t x[4], y[4];
    std::memcpy(x, a, 32);
    std::memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
           __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

void match(const uint8_t *d1, int n1, const uint8_t *d2, int n2, int32_t *idx, int32_t *dist,
           int32_t *second, int32_t *rev_idx) {
    std::vector<int> rev_best;
    if (rev_idx) {
        rev_best.assign(n2, INT_MAX);
        for (int j = 0; j < n2; j++) rev_idx[j] = -1;
    }
    for (int i = 0; i < n1; i++) {  // src/BriefDescriptor.cc:165-181
        int best = INT_MAX, sec = INT_MAX, bj = -1;
        for (int j = 0; j < n2; j++) {
            int d = hamming_fast(d1 + (size_t)i * 32, d2 + (size_t)j * 32);
            if (d < best) {  // strict: lowest j among equal minima
                sec = best;
                best = d;
                bj = j;
            } else if (d < sec) {
                sec = d;
            }
            if (rev_idx && d < rev_best[j]) {
                rev_best[j] = d;
                rev_idx[j] = i;
            }
        }
        idx[i] = bj;
        dist[i] = best;
        if (second) second[i] = sec;
    }
}

// ---- transparent introsort (libstdc++ bits/stl_algo.h: __introsort_loop, __unguarded_partition_pivot,
// __move_median_to_first, __final_insertion_sort), pruned to what decides positions [0,k) ----
struct SP {
    float s;
    int32_t p;
};
inline bool before(const SP &a, const SP &b) { return a.s > b.s; }  // src/FastDetector.cc:343-345

void introsort_loop_topk(SP *A, int first, int last, int depth, int k) {
    while (last - first > 16) {
        if (first >= k) return;  // nothing in this range can reach a position < k
        if (depth == 0) {
            std::make_heap(A + first, A + last, before);  // __partial_sort(first,last,last)
            std::sort_heap(A + first, A + last, before);
            return;
        }
        --depth;
        int mid = first + (last - first) / 2;
        int a = first + 1, b = mid, c = last - 1, res = first;
        if (before(A[a], A[b])) {
            if (before(A[b], A[c])) std::swap(A[res], A[b]);
            else if (before(A[a], A[c])) std::swap(A[res], A[c]);
            else std::swap(A[res], A[a]);
        } else if (before(A[a], A[c])) std::swap(A[res], A[a]);
        else if (before(A[b], A[c])) std::swap(A[res], A[c]);
        else std::swap(A[res], A[b]);
        int f = first + 1, l = last;
        const SP piv = A[first];
        while (true) {
            while (before(A[f], piv)) ++f;
            --l;
            while (before(piv, A[l])) --l;
            if (!(f < l)) break;
            std::swap(A[f], A[l]);
            ++f;
        }
        int cut = f;
        introsort_loop_topk(A, cut, last, depth, k);
        last = cut;
    }
}

}  // namespace

extern "C" {

void yavo_oracle_ring(int xc, int yc, int32_t *out_xy) {
    for (int k = 0; k < 16; k++) {
        out_xy[2 * k] = xc + kRing[k][0];
        out_xy[2 * k + 1] = yc + kRing[k][1];
    }
}

int yavo_oracle_ring_literal(int xc, int yc, int32_t *out_xy) {
    // Follows src/FastDetector.cc:50-112 step by step, including the two std::set
    // containers whose comparators (:19-45) compare the second coordinate only.
    const int R = 3;
    typedef std::pair<int, int> P;
    auto cmpA = [](const P &a, const P &b) { return a.second < b.second; };
    auto cmpB = [](const P &a, const P &b) { return a.second > b.second; };
    std::set<P, decltype(cmpA)> firstHalf(cmpA);
    std::set<P, decltype(cmpB)> secondHalf(cmpB);
    int xl = 0, yl = R, d = 3 - 2 * R;
    while (yl >= xl) {
        xl++;
        if (d <= 0) {
            d = d + 4 * xl + 6;
        } else {
            yl--;
            d = d + 4 * (xl - yl) + 10;
        }
        const int sym[8][2] = {{xl, yl}, {yl, xl}, {yl, -xl}, {xl, -yl},
                               {-xl, -yl}, {-yl, -xl}, {-yl, xl}, {-xl, yl}};  // :114-116
        for (int i = 0; i < 8; i++) {
            int sx = sym[i][0], sy = sym[i][1];
            int xa = (sx >= 0) ? xc + std::abs(sx) : xc - std::abs(sx);
            int ya = (sy >= 0) ? yc - std::abs(sy) : yc + std::abs(sy);
            if (sx >= 0) firstHalf.insert(P(xa, ya));
            else secondHalf.insert(P(xa, ya));
        }
    }
    firstHalf.insert(P(xc + R, yc));
    secondHalf.insert(P(xc - R, yc));
    std::vector<P> pts;
    pts.push_back(P(xc, yc - R));
    for (auto &e : firstHalf) pts.push_back(e);
    pts.push_back(P(xc, yc + R));
    for (auto &e : secondHalf) pts.push_back(e);
    int n = (int)pts.size();
    for (int k = 0; k < n && k < 16; k++) {
        out_xy[2 * k] = pts[k].first;
        out_xy[2 * k + 1] = pts[k].second;
    }
    return n;
}

int yavo_oracle_check_in_between(uint8_t cent, uint8_t cond) { return in_between(cent, cond) ? 1 : 0; }

int yavo_oracle_check_contiguous(uint8_t cent, const uint8_t ring_vals[16]) {
    int run = 0;
    for (int k = 0; k < 16; k++) {
        if (in_between(cent, ring_vals[k])) run = 0;
        else run++;
        if (run >= kRunLength) return 1;
    }
    return 0;
}

void yavo_oracle_sobel(const uint8_t *img, int H, int W, float *Ix, float *Iy) {
    // src/FastDetector.cc:164-200: zero border of 1, outputs written for r in [0,H-3], c in [0,W-3]
    std::memset(Ix, 0, sizeof(float) * (size_t)H * W);
    std::memset(Iy, 0, sizeof(float) * (size_t)H * W);
    auto px = [&](int r, int c) -> float {
        return (r < 0 || c < 0 || r >= H || c >= W) ? 0.0f : (float)img[(size_t)r * W + c];
    };
    const float kx[3][3] = {{-1, 0, 1}, {-2, 0, 2}, {-1, 0, 1}};
    const float ky[3][3] = {{-1, -2, -1}, {0, 0, 0}, {1, 2, 1}};
    for (int r = 0; r <= H - 3; r++)
        for (int c = 0; c <= W - 3; c++) {
            float sx = 0, sy = 0;
            for (int k = 0; k < 3; k++)
                for (int l = 0; l < 3; l++) {
                    float v = px(r + k - 1, c + l - 1);
                    sx += kx[k][l] * v;
                    sy += ky[k][l] * v;
                }
            Ix[(size_t)r * W + c] = sx;
            Iy[(size_t)r * W + c] = sy;
        }
}

void yavo_oracle_eigen2x2(float a, float b, float c, float *l1, float *l2) { eigen2x2(a, b, c, l1, l2); }

float yavo_oracle_score_from_tensor(float a, float b, float c) {
    float l1, l2;
    eigen2x2(a, b, c, &l1, &l2);
    return harris_from_eigen(l1, l2);
}

float yavo_oracle_harris(const uint8_t *img, int H, int W, int x, int y) {
    // general version through the Sobel planes (handles the zero border / zero tail rows)
    std::vector<float> Ix((size_t)H * W), Iy((size_t)H * W);
    yavo_oracle_sobel(img, H, W, Ix.data(), Iy.data());
    float m00 = 0, m01 = 0, m11 = 0;
    for (int i = x - 1; i <= x + 1; i++)
        for (int j = y - 1; j <= y + 1; j++) {
            float gx = Ix[(size_t)i * W + j], gy = Iy[(size_t)i * W + j];
            m00 += gx * gx;
            m01 += gx * gy;
            m11 += gy * gy;
        }
    float l1, l2;
    eigen2x2(m00, m01, m11, &l1, &l2);
    return harris_from_eigen(l1, l2);
}

int yavo_oracle_fast_candidates(const uint8_t *img, int H, int W, int cap, int32_t *rows,
                                int32_t *cols, float *scores) {
    std::vector<Cand> c;
    collect_candidates(img, H, W, c);
    int n = (int)c.size();
    for (int i = 0; i < n && i < cap; i++) {
        rows[i] = c[i].x;
        cols[i] = c[i].y;
        scores[i] = c[i].score;
    }
    return n;
}

int yavo_oracle_fast_detect(const uint8_t *img, int H, int W, int max_kp, int32_t *rows,
                            int32_t *cols, float *scores, int *n_cand) {
    std::vector<Cand> c;
    collect_candidates(img, H, W, c);
    if (n_cand) *n_cand = (int)c.size();
    std::sort(c.begin(), c.end(), [](const Cand &a, const Cand &b) { return a.score > b.score; });
    int n = std::min((int)c.size(), max_kp);
    for (int i = 0; i < n; i++) {
        rows[i] = c[i].x;
        cols[i] = c[i].y;
        if (scores) scores[i] = c[i].score;
    }
    return n;
}

void yavo_oracle_std_sort_desc(float *scores, int32_t *payload, int n) {
    // element type mirrors the reference's FastFeature {int x; int y; float} only in that
    // std::sort's control flow depends on comparisons alone, not on the element size
    std::vector<SP> v(n);
    for (int i = 0; i < n; i++) v[i] = {scores[i], payload[i]};
    std::sort(v.begin(), v.end(), before);
    for (int i = 0; i < n; i++) {
        scores[i] = v[i].s;
        payload[i] = v[i].p;
    }
}

void yavo_oracle_introsort_topk(float *scores, int32_t *payload, int n, int k) {
    yavo_oracle_introsort_topk_depth(scores, payload, n, k, -1);
}

void yavo_oracle_introsort_topk_depth(float *scores, int32_t *payload, int n, int k, int depth) {
    if (n == 0) return;
    std::vector<SP> v(n);
    for (int i = 0; i < n; i++) v[i] = {scores[i], payload[i]};
    int lg = 31 - __builtin_clz((unsigned)n);
    introsort_loop_topk(v.data(), 0, n, depth >= 0 ? depth : 2 * lg, k);
    // __final_insertion_sort == stable insertion sort; only the prefix that can reach [0,k) matters
    int lim = std::min(n, k + 16);
    for (int i = 1; i < lim; i++) {
        SP val = v[i];
        int j = i;
        while (j > 0 && before(val, v[j - 1])) {
            v[j] = v[j - 1];
            j--;
        }
        v[j] = val;
    }
    for (int i = 0; i < n; i++) {
        scores[i] = v[i].s;
        payload[i] = v[i].p;
    }
}

void yavo_oracle_gaussian_blur(const uint8_t *img, int H, int W, uint8_t *out) { gaussian_blur(img, H, W, out); }

void yavo_oracle_brief(const uint8_t *img, const uint8_t *blurred, int H, int W, const int32_t *offsets,
                       const int32_t *rows, const int32_t *cols, int n, uint8_t *desc, uint8_t *valid,
                       int *n_oob) {
    std::vector<uint8_t> tmp;
    if (!blurred) {
        tmp.resize((size_t)H * W);
        gaussian_blur(img, H, W, tmp.data());  // src/BriefDescriptor.cc:90
        blurred = tmp.data();
    }
    brief(blurred, H, W, offsets, rows, cols, n, desc, valid, n_oob);
}

int yavo_oracle_popcount(uint8_t v) { return popcount_serial(v); }

int yavo_oracle_hamming(const uint8_t *a, const uint8_t *b) {  // src/BriefDescriptor.cc:139-146
    int d = 0;
    for (int i = 0; i < 32; i++) d += popcount_serial((uint8_t)(a[i] ^ b[i]));
    return d;
}

void yavo_oracle_match(const uint8_t *d1, int n1, const uint8_t *d2, int n2, int32_t *idx, int32_t *dist,
                       int32_t *second, int32_t *rev_idx) {
    match(d1, n1, d2, n2, idx, dist, second, rev_idx);
}

int yavo_oracle_remove_outliers(const int32_t *dist, int n, int threshold, uint8_t *keep) {
    if (n == 0) return 0;  // the reference dereferences end() here (UB); defined as "nothing kept"
    int mn = INT_MAX;
    for (int i = 0; i < n; i++) mn = std::min(mn, dist[i]);
    // 2 * distance in the reference's int arithmetic: with an empty train set every distance is INT_MAX and the
    // product wraps to -2 on the reference's targets (signed overflow, UB by the letter), so nothing is kept
    const int lim = std::max((int)(2u * (unsigned)mn), threshold);
    int kept = 0;
    for (int i = 0; i < n; i++) {
        keep[i] = (dist[i] < lim) ? 1 : 0;
        kept += keep[i];
    }
    return kept;
}

int yavo_oracle_pipeline(const uint8_t *frames, int F, int H, int W, const int32_t *offsets, int max_kp,
                         int do_match, int nthreads, int32_t *kp_rows, int32_t *kp_cols, float *kp_scores,
                         uint8_t *kp_desc, int32_t *n_kp, int32_t *match_idx, int32_t *match_dist) {
    if (nthreads < 1) nthreads = 1;
    const size_t fsz = (size_t)H * W;
    // per-frame results are kept so that the match of (f-1, f) can run after both exist
    std::vector<std::vector<uint8_t>> descs(F);
    std::vector<int> counts(F, 0);
    std::atomic<int> next(0);
    auto detect_worker = [&]() {
        std::vector<int32_t> r(max_kp), c(max_kp);
        std::vector<float> s(max_kp);
        std::vector<uint8_t> d((size_t)max_kp * 32), v(max_kp);
        for (;;) {
            int f = next.fetch_add(1);
            if (f >= F) break;
            const uint8_t *img = frames + fsz * f;
            int nc = 0;
            int n = yavo_oracle_fast_detect(img, H, W, max_kp, r.data(), c.data(), s.data(), &nc);
            int oob = 0;
            yavo_oracle_brief(img, nullptr, H, W, offsets, r.data(), c.data(), n, d.data(), v.data(), &oob);
            // Brief::computeBrief keeps only keypoints passing checkBoundry (src/BriefDescriptor.cc:97,121)
            int m = 0;
            descs[f].resize((size_t)n * 32);
            for (int i = 0; i < n; i++) {
                if (!v[i]) continue;
                std::memcpy(descs[f].data() + (size_t)m * 32, d.data() + (size_t)i * 32, 32);
                if (kp_rows) kp_rows[(size_t)f * max_kp + m] = r[i];
                if (kp_cols) kp_cols[(size_t)f * max_kp + m] = c[i];
                if (kp_scores) kp_scores[(size_t)f * max_kp + m] = s[i];
                if (kp_desc) std::memcpy(kp_desc + ((size_t)f * max_kp + m) * 32, d.data() + (size_t)i * 32, 32);
                m++;
            }
            descs[f].resize((size_t)m * 32);
            counts[f] = m;
            if (n_kp) n_kp[f] = m;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(detect_worker);
        for (auto &t : th) t.join();
    }
    if (do_match) {
        next = 1;
        auto match_worker = [&]() {
            std::vector<int32_t> mi(max_kp), md(max_kp);
            for (;;) {
                int f = next.fetch_add(1);
                if (f >= F) break;
                int n1 = counts[f - 1], n2 = counts[f];
                match(descs[f - 1].data(), n1, descs[f].data(), n2, mi.data(), md.data(), nullptr, nullptr);
                if (match_idx) std::memcpy(match_idx + (size_t)f * max_kp, mi.data(), sizeof(int32_t) * n1);
                if (match_dist) std::memcpy(match_dist + (size_t)f * max_kp, md.data(), sizeof(int32_t) * n1);
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(match_worker);
        for (auto &t : th) t.join();
    }
    return 0;
}

}  // extern "C"
