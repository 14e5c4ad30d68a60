/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Plain-C restatement of the reference's feature extraction
 *   LaserProcessingClass::featureExtraction           /root/reference/src/laserProcessingClass.cpp:10-96
 *   LaserProcessingClass::featureExtractionFromSector /root/reference/src/laserProcessingClass.cpp:99-209
 * Pinned against the real reference source compiled in place (oracle/_ref/libpf_ref_extract.so,
 * see oracle/Makefile and tests/test_oracle_extract.py).
 *
 * Conventions the reference leaves open (std::sort is unstable, :101-104): equal curvature values
 * are ordered by lower ring position first.
 *
 * order = 0: reference emission order (surf ascending by curvature inside a sector, :198-205)
 * order = 1: product emission order (surf ascending by ring position inside a sector); the SETS are identical.
 * Compile with -ffp-contract=off: the float stencil at :74-76 must not be fused.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int id; double value; } curv_t;

static int curv_less(const void* a, const void* b) {
    const curv_t* x = (const curv_t*)a; const curv_t* y = (const curv_t*)b;
    if (x->value < y->value) return -1;
    if (x->value > y->value) return 1;
    return (x->id > y->id) - (x->id < y->id);
}

/* ring id of one point, or -1 when the reference drops it (:25-61) */
int pforacle_ring_id(float x, float y, float z, int num_lines, double min_distance, double max_distance) {
    float s = x * x + y * y;                 /* float arithmetic (:25) */
    double distance = (double)sqrtf(s);      /* sqrt(float) resolves to the float overload, see oracle/shim/pcl/point_types.h */
    if (distance < min_distance || distance > max_distance) return -1;
    double angle = atan((double)z / distance) * 180 / M_PI;
    int id;
    if (num_lines == 16) {
        id = (int)((angle + 15) / 2 + 0.5);
        if (id > 15 || id < 0) return -1;
    } else if (num_lines == 32) {
        id = (int)((angle + 92.0 / 3.0) * 3.0 / 4.0);
        if (id > 31 || id < 0) return -1;
    } else if (num_lines == 64) {
        if (angle >= -8.83) id = (int)((2 - angle) * 3.0 + 0.5);
        else id = 32 + (int)((-8.83 - angle) * 2.0 + 0.5);
        if (angle > 2 || angle < -24.33 || id > 63 || id < 0) return -1;
    } else {
        return -1;   /* reference prints "wrong scan number" and files the point under ring 0; we reject the config instead */
    }
    return id;
}

static double step_d2(const float* p, int a, int b) {   /* (:129-132): float differences, double squares */
    double dx = p[4 * a + 0] - p[4 * b + 0];
    double dy = p[4 * a + 1] - p[4 * b + 1];
    double dz = p[4 * a + 2] - p[4 * b + 2];
    return dx * dx + dy * dy + dz * dz;
}

/* xyzi: n points of 4 floats.  Outputs are INPUT INDICES in emission order; label[i] = 0 none, 1 edge, 2 surf.
 * ring_of (optional, n ints) receives the ring id of every input point.  Returns 0. */
int pforacle_extract(const float* xyzi, int n, int num_lines, double min_distance, double max_distance, int order,
                     int* edge_idx, int* n_edge, int* surf_idx, int* n_surf, unsigned char* label, int* ring_of) {
    int* ring = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
    int* count = (int*)calloc(64, sizeof(int));
    int ne = 0, ns = 0;
    if (label) memset(label, 0, n);
    for (int i = 0; i < n; ++i) {
        ring[i] = pforacle_ring_id(xyzi[4 * i], xyzi[4 * i + 1], xyzi[4 * i + 2], num_lines, min_distance, max_distance);
        if (ring_of) ring_of[i] = ring[i];
        if (ring[i] >= 0) count[ring[i]]++;
    }
    for (int r = 0; r < num_lines; ++r) {
        int m = count[r];
        if (m < 131) continue;                                            /* (:67) */
        int* src = (int*)malloc(sizeof(int) * m);
        float* p = (float*)malloc(sizeof(float) * 4 * m);
        int k = 0;
        for (int i = 0; i < n; ++i)
            if (ring[i] == r) { src[k] = i; memcpy(p + 4 * k, xyzi + 4 * i, 16); ++k; }   /* stable (:62) */
        int total = m - 10;                                               /* (:72) */
        curv_t* c = (curv_t*)malloc(sizeof(curv_t) * total);
        for (int j = 5; j < m - 5; ++j) {                                 /* (:73-80) float sums, left to right */
            float d[3];
            for (int a = 0; a < 3; ++a) {
                float t = p[4 * (j - 5) + a] + p[4 * (j - 4) + a];
                t = t + p[4 * (j - 3) + a];
                t = t + p[4 * (j - 2) + a];
                t = t + p[4 * (j - 1) + a];
                t = t - 10 * p[4 * j + a];
                t = t + p[4 * (j + 1) + a];
                t = t + p[4 * (j + 2) + a];
                t = t + p[4 * (j + 3) + a];
                t = t + p[4 * (j + 4) + a];
                t = t + p[4 * (j + 5) + a];
                d[a] = t;
            }
            double dx = d[0], dy = d[1], dz = d[2];
            c[j - 5].id = j;
            c[j - 5].value = dx * dx + dy * dy + dz * dz;
        }
        unsigned char* picked = (unsigned char*)malloc(m);
        curv_t* sec = (curv_t*)malloc(sizeof(curv_t) * (total + 1));
        int L = total / 6;                                                /* (:82) */
        for (int s = 0; s < 6; ++s) {
            int lo = L * s, hi = (s == 5) ? total - 1 : L * (s + 1) - 1;  /* hi excluded (:83-88) */
            int len = hi - lo;
            if (len < 0) len = 0;
            memcpy(sec, c + lo, sizeof(curv_t) * len);
            qsort(sec, len, sizeof(curv_t), curv_less);                   /* (:101-104), ties by lower id */
            memset(picked, 0, m);
            int cnt = 0;
            for (int i = len - 1; i >= 0; --i) {                          /* (:110-148) */
                int ind = sec[i].id;
                if (picked[ind]) continue;
                if (sec[i].value <= 0.1) break;
                cnt++;
                picked[ind] = 1;
                if (cnt <= 20) {
                    edge_idx[ne++] = src[ind];
                    if (label) label[src[ind]] = 1;
                } else {
                    break;
                }
                for (int q = 1; q <= 5; ++q) {
                    if (step_d2(p, ind + q, ind + q - 1) > 0.05) break;
                    picked[ind + q] = 1;
                }
                for (int q = -1; q >= -5; --q) {
                    if (step_d2(p, ind + q, ind + q + 1) > 0.05) break;
                    picked[ind + q] = 1;
                }
            }
            if (order == 0) {
                for (int i = 0; i < len; ++i) {                           /* (:198-205) */
                    int ind = sec[i].id;
                    if (!picked[ind]) { surf_idx[ns++] = src[ind]; if (label) label[src[ind]] = 2; }
                }
            } else {
                for (int i = 0; i < len; ++i) {
                    int ind = c[lo + i].id;
                    if (!picked[ind]) { surf_idx[ns++] = src[ind]; if (label) label[src[ind]] = 2; }
                }
            }
        }
        free(sec); free(picked); free(c); free(p); free(src);
    }
    *n_edge = ne; *n_surf = ns;
    free(count); free(ring);
    return 0;
}
