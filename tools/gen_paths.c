/*
 * gen_paths.c -- seeded synthetic path workloads for the BASELINE.json configs (inputs only).
 *
 * The reference ships no workload generator; these follow SURVEY.md section 8d:
 *   G3  glyph atlas   : quadratic TrueType-style outlines, 12-48 px, y-flipped font transform
 *   G4  stress        : closed cubic blobs (3-64 segments) on a 4096x4096 canvas
 *   G5a giant         : ONE path of concentric cubic rings on a 16384x16384 canvas
 * Path i of G3/G4 is a pure function of (kind, i): SplitMix64 stream seeded with
 * base + i * 0x9E3779B97F4A7C15, uniform u = (next() >> 40) * 2^-24.  Geometry is computed in
 * double and cast to float at the end, so every consumer (oracle, GPU, every rank of a
 * multi-GPU run) sees byte-identical OchreCmd arrays.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { uint32_t tag; float v[6]; } OchreCmd;
enum { MOVE = 0, LINE = 1, QUAD = 2, CUBIC = 3, CONIC = 4, CLOSE = 5 };

typedef struct { uint64_t s; } Rng;
static inline uint64_t rng_next(Rng *r) {
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double rng_u(Rng *r) { return (double)(rng_next(r) >> 40) * (1.0 / 16777216.0); }
static inline Rng rng_for(uint64_t base, uint64_t i) { Rng r; r.s = base + i * 0x9E3779B97F4A7C15ull; return r; }

#define TWO_PI 6.283185307179586476925286766559

typedef struct { OchreCmd *out; uint64_t n; } Sink; /* out == NULL: count only */
static inline void put(Sink *s, uint32_t tag, double a, double b, double c, double d, double e, double f) {
    if (s->out) {
        OchreCmd *q = &s->out[s->n];
        q->tag = tag;
        q->v[0] = (float)a; q->v[1] = (float)b; q->v[2] = (float)c; q->v[3] = (float)d; q->v[4] = (float)e; q->v[5] = (float)f;
    }
    s->n++;
}

/* ---- G4: cubic blobs ------------------------------------------------------------------- */
static inline double clampc(double v, double hi) { return v < 0.0 ? 0.0 : (v > hi ? hi : v); }
static void g4_path(uint64_t i, Sink *s, float *xf) {
    Rng r = rng_for(0x0C4E0004ull, i);
    int n = 3 + (int)(rng_u(&r) * 62.0);
    if (n > 64) n = 64;
    double cx = rng_u(&r) * 4096.0, cy = rng_u(&r) * 4096.0;
    double R = 8.0 * pow(2.0, 5.0 * rng_u(&r));
    double phi = rng_u(&r);
    const double hi = 4096.0 - 1.0 / 64.0;
    double vx[64], vy[64];
    for (int m = 0; m < n; m++) {
        double rad = R * (0.4 + 0.6 * rng_u(&r));
        double ang = TWO_PI * ((double)m + phi) / (double)n;
        vx[m] = clampc(cx + rad * cos(ang), hi);
        vy[m] = clampc(cy + rad * sin(ang), hi);
    }
    put(s, MOVE, vx[0], vy[0], 0, 0, 0, 0);
    for (int m = 0; m < n; m++) {
        int m1 = (m + 1) % n;
        double c1x = clampc(vx[m] + R * (rng_u(&r) - 0.5), hi), c1y = clampc(vy[m] + R * (rng_u(&r) - 0.5), hi);
        double c2x = clampc(vx[m1] + R * (rng_u(&r) - 0.5), hi), c2y = clampc(vy[m1] + R * (rng_u(&r) - 0.5), hi);
        put(s, CUBIC, c1x, c1y, c2x, c2y, vx[m1], vy[m1]);
    }
    put(s, CLOSE, 0, 0, 0, 0, 0, 0);
    if (xf) { xf[0] = 1; xf[1] = 0; xf[2] = 0; xf[3] = 1; xf[4] = 0; xf[5] = 0; }
}

/* ---- G3: glyph outlines ---------------------------------------------------------------- */
static void g3_path(uint64_t i, Sink *s, float *xf) {
    Rng r = rng_for(0x0C4E0003ull, i);
    double px = 12.0 + 36.0 * rng_u(&r);
    double ox = rng_u(&r), oy = rng_u(&r);
    int k = 1 + (int)(rng_u(&r) * 3.0);
    if (k > 3) k = 3;
    static const double scale[3] = {900.0, 420.0, 180.0};
    for (int j = 0; j < k; j++) {
        double R = scale[j];
        int n = 4 + (int)(rng_u(&r) * 9.0);
        if (n > 12) n = 12;
        double phi = rng_u(&r);
        double onx[12], ony[12], offx[12], offy[12];
        int is_line[12];
        for (int m = 0; m < n; m++) {
            double rad = R * (0.70 + 0.30 * rng_u(&r));
            double ang = TWO_PI * ((double)m + phi) / (double)n;
            onx[m] = floor(1024.0 + rad * cos(ang) + 0.5);
            ony[m] = floor(1024.0 + rad * sin(ang) + 0.5);
            double rad2 = R * (0.75 + 0.45 * rng_u(&r));
            double ang2 = TWO_PI * ((double)m + 0.5 + phi) / (double)n;
            offx[m] = floor(1024.0 + rad2 * cos(ang2) + 0.5); /* off-curve point of edge m -> m+1 */
            offy[m] = floor(1024.0 + rad2 * sin(ang2) + 0.5);
            is_line[m] = rng_u(&r) < 0.25;
        }
        if ((j & 1) == 0) { /* outer contours forward */
            put(s, MOVE, onx[0], ony[0], 0, 0, 0, 0);
            for (int m = 0; m < n; m++) {
                int m1 = (m + 1) % n;
                if (is_line[m]) put(s, LINE, onx[m1], ony[m1], 0, 0, 0, 0);
                else put(s, QUAD, offx[m], offy[m], onx[m1], ony[m1], 0, 0);
            }
        } else { /* inner contours reversed */
            put(s, MOVE, onx[0], ony[0], 0, 0, 0, 0);
            for (int m = n - 1; m >= 0; m--) { /* edge m (m -> m+1) walked backwards ends at on[m] */
                if (is_line[m]) put(s, LINE, onx[m], ony[m], 0, 0, 0, 0);
                else put(s, QUAD, offx[m], offy[m], onx[m], ony[m], 0, 0);
            }
        }
        put(s, CLOSE, 0, 0, 0, 0, 0, 0);
    }
    if (xf) {
        double sc = px / 2048.0;
        xf[0] = (float)sc; xf[1] = 0; xf[2] = 0; xf[3] = (float)-sc; xf[4] = (float)ox; xf[5] = (float)(px + oy);
    }
}

/* ---- G5a: one path of concentric rings ------------------------------------------------- */
/* rings: number of rings; spacing: radial spacing in px; segs: cubic segments per ring */
static void g5a_path(uint32_t rings, double spacing, uint32_t segs, Sink *s, float *xf) {
    Rng r = rng_for(0x0C4E0005ull, 0);
    const double cx = 8192.0, cy = 8192.0;
    for (uint32_t k = 1; k <= rings; k++) {
        double base = spacing * (double)k;
        double dir = (k & 1) ? 1.0 : -1.0;
        double dth = dir * TWO_PI / (double)segs;
        double h = 4.0 / 3.0 * tan(fabs(dth) / 4.0);
        double rad_prev = base + 4.0 * (rng_u(&r) - 0.5);
        double rad0 = rad_prev;
        double th = TWO_PI * rng_u(&r);
        double x0 = cx + rad_prev * cos(th), y0 = cy + rad_prev * sin(th);
        put(s, MOVE, x0, y0, 0, 0, 0, 0);
        for (uint32_t m = 0; m < segs; m++) {
            double th1 = th + dth;
            double rad = (m + 1 == segs) ? rad0 : base + 4.0 * (rng_u(&r) - 0.5);
            double x1 = cx + rad * cos(th1), y1 = cy + rad * sin(th1);
            /* tangent handles of a circular arc, scaled by each end's radius */
            double c1x = x0 - dir * h * rad_prev * sin(th), c1y = y0 + dir * h * rad_prev * cos(th);
            double c2x = x1 + dir * h * rad * sin(th1), c2y = y1 - dir * h * rad * cos(th1);
            put(s, CUBIC, c1x, c1y, c2x, c2y, x1, y1);
            th = th1; rad_prev = rad; x0 = x1; y0 = y1;
        }
        put(s, CLOSE, 0, 0, 0, 0, 0, 0);
    }
    if (xf) { xf[0] = 1; xf[1] = 0; xf[2] = 0; xf[3] = 1; xf[4] = 0; xf[5] = 0; }
}

/* kind: 3 = G3, 4 = G4.  Fills cmd_off[0..n] (relative to this call) and, when cmds != NULL, cmds and xf.
 * Returns the number of commands. */
__attribute__((visibility("default")))
uint64_t gen_paths(int kind, uint64_t first, uint32_t n, OchreCmd *cmds, uint32_t *cmd_off, float *xf) {
    Sink s = {cmds, 0};
    for (uint32_t i = 0; i < n; i++) {
        if (cmd_off) cmd_off[i] = (uint32_t)s.n;
        float *x = (cmds && xf) ? xf + 6 * (uint64_t)i : 0;
        if (kind == 3) g3_path(first + i, &s, x); else g4_path(first + i, &s, x);
    }
    if (cmd_off) cmd_off[n] = (uint32_t)s.n;
    return s.n;
}

__attribute__((visibility("default")))
uint64_t gen_rings(uint32_t rings, double spacing, uint32_t segs, OchreCmd *cmds, float *xf) {
    Sink s = {cmds, 0};
    g5a_path(rings, spacing, segs, &s, cmds ? xf : 0);
    return s.n;
}
