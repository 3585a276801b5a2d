// Polygon -> polar ray targets for one anchor per thread, list-free ("monotone arcs") form.
//
// Reference semantics (utils/tal.py:1257-1277 == 1172-1193): for each of R fixed angles take the 4 contour
// points nearest in angle, target = max of their distances, 1e-6 when the nearest is more than 3 degrees away.
//
// Seen from the anchor, the closed contour is a chain of angularly monotone arcs separated by reversal
// points (local extrema of the angle; none at all when the anchor sees the contour star-shaped).  Along a
// monotone arc the points nearest to a ray are the contour indices around the step that crosses it, so:
//   sweep    one pass over the 360 points (broadcast from shared memory), four points per iteration: the
//            direction of every step (sign of the cross product of consecutive view vectors), the angular
//            sector (between ray s and ray s+1) of the group's last point; every ray crossed by a group is
//            recorded with the index of the crossing step (up to 4 crossings per ray, 2 bytes each); a
//            group that contains a direction change is only noted (reversal list);
//   phase C  the noted groups are re-walked point by point: exact crossings, the reversal points, and for
//            every reversal a pseudo-crossing of the one neighbouring ray the arc did not reach;
//   pass 1   per ray: the usual case - one crossing - is settled from the six points around it when the
//            two outer ones are not nearer than the inner four (then no other point of the arc is);
//   pass 2   everything else (several crossings, reversal neighbourhoods, lopsided spacing) goes through a
//            per-warp queue: all windows of the ray are merged into one top-4 list and grown until their
//            border points are farther than the 4th key; rays whose 4th key exceeds one ray spacing also
//            take the neighbourhoods of all reversal points.  What cannot be handled (more than 4 crossings
//            of a ray, more than 8 reversals, a contour point on the anchor) is scanned exactly.
// Why this is exact: an unevaluated point q can be walked along its arc towards the ray with decreasing
// angular distance until it meets a recorded crossing or a reversal point; the border point of that
// window lies on the way, so key(q) >= key(border) >= 4th key.  A reversal whose pseudo-crossing belongs
// to another ray is more than one ray spacing away from this ray.
//
// The file compiles for the device (nvcc) and for the host (g++, tests/test_arcs_host.py drives it against
// the oracle); everything per-thread is YA_HD, warp-level code is device only.
#pragma once
#include <stdint.h>
#include <math.h>
#include <cuda_runtime.h>

#ifdef __CUDACC__
#define YA_HD __host__ __device__ __forceinline__
#else
#define YA_HD inline
#endif

#ifndef YCR_C
#define YCR_C 360
#endif
#define YA_KMAX 4          // crossings kept per ray
#define YA_RMAX 12         // reversal groups / reversal points kept per candidate
#define YA_PSEUDO 0x200u   // slot flag: reversal neighbourhood, not a crossing
#define YA_FLOOR 1e-6f
#define YA_GATE_DEG 3.0
#define YA_TOL_DEG 0.01    // slack of the "one ray spacing" certificate
#define YA_GROW 24         // window growth limit per side (then: exact scan)
#define YA_PAD 4           // contour copy is padded by 4 wrapped points on both sides

struct ArcConst {
    float g2;      // sin^2(3 deg): nearest key above this -> nothing within the gate
    float kstep;   // sin^2(360/R - TOL): 4th key below this -> reversals of other rays cannot matter
};

static inline ArcConst make_arc_const(int R) {
    ArcConst c;
    const double d2r = 3.14159265358979323846 / 180.0;
    const double sg = sin(YA_GATE_DEG * d2r), ss = sin((360.0 / R - YA_TOL_DEG) * d2r);
    c.g2 = (float)(sg * sg);
    c.kstep = (float)(ss * ss);
    return c;
}

template <int R, int NT>
struct alignas(16) ArcSmem {
    ushort4 slot[R][NT];                 // crossings of ray i: index | flags; the first 4 bytes become the target
    float2 cpad[YCR_C + 2 * YA_PAD];     // contour, cpad[k] = point (k - YA_PAD) mod 360
    float2 raydir[R + 1];                // (cos, sin) of i*360/R deg; entry R repeats entry 0
    float2 anchor[NT];
    unsigned short rev[YA_RMAX][NT];     // sweep: group | sector << 7 | (previous direction negative) << 14
    unsigned short rpt[YA_RMAX][NT];     // phase C: reversal point indices
    unsigned short queue[NT / 32 > 0 ? NT / 32 : 1][R * 32];   // per warp: (thread << 7) | ray
    unsigned char cnt[R][NT];            // crossings recorded per ray (may exceed YA_KMAX: then the ray is scanned)
    unsigned char nrev[NT], nrpt[NT], bad[NT];
    YA_HD float& tv(int i, int t) { return reinterpret_cast<float*>(&slot[i][t])[0]; }
    YA_HD float tv(int i, int t) const { return reinterpret_cast<const float*>(&slot[i][t])[0]; }
};

YA_HD float ya_cross(float ax, float ay, float bx, float by) { return fmaf(ax, by, -(ay * bx)); }

YA_HD float ya_rcp(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

YA_HD uint32_t ya_bits(float x) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    union { float f; uint32_t u; } v;
    v.f = x;
    return v.u;
#endif
}

YA_HD int ya_wrap(int j) { return (j < 0) ? j + YCR_C : ((j >= YCR_C) ? j - YCR_C : j); }

// sector s with cross(raydir[s], v) >= 0 and cross(raydir[s+1], v) < 0
template <int R, int NT>
YA_HD int arc_sector_of(const ArcSmem<R, NT>& sm, float vx, float vy) {
    float t = atan2f(vy, vx) * (float)(R / 6.283185307179586);
    if (t < 0.f) t += (float)R;
    int s = (int)t;
    s = (s < 0) ? 0 : ((s >= R) ? R - 1 : s);
    for (int guard = 0; guard < R; ++guard) {
        const float2 lo = sm.raydir[s], hi = sm.raydir[s + 1];
        if (ya_cross(lo.x, lo.y, vx, vy) < 0.f) s = (s == 0) ? R - 1 : s - 1;
        else if (ya_cross(hi.x, hi.y, vx, vy) >= 0.f) s = (s == R - 1) ? 0 : s + 1;
        else break;
    }
    return s;
}

template <int R, int NT>
YA_HD void arc_log(ArcSmem<R, NT>& sm, int tid, int ray, int c, unsigned flags) {
    const int n = sm.cnt[ray][tid];
    if (n < YA_KMAX) reinterpret_cast<unsigned short*>(&sm.slot[ray][tid])[n] = (unsigned short)((unsigned)c | flags);
    sm.cnt[ray][tid] = (unsigned char)((n < 255) ? n + 1 : 255);
}

// ------------------------------------------------------------------------------------------------
// sweep
// ------------------------------------------------------------------------------------------------
template <int R, int NT>
YA_HD void arc_sweep(ArcSmem<R, NT>& sm, int tid, float ax, float ay, float near2) {
    for (int i = 0; i < R; ++i) sm.cnt[i][tid] = 0;
    const float2* P = sm.cpad + YA_PAD;
    float vpx = P[-1].x - ax, vpy = P[-1].y - ay;
    float pdb = ya_cross(P[-2].x - ax, P[-2].y - ay, vpx, vpy);   // direction of the step into point 359
    int s = arc_sector_of<R, NT>(sm, vpx, vpy);
    int nrev = 0;
    bool bad = false;
    for (int g = 0; g < YCR_C / 4; ++g) {
        const float4 q01 = reinterpret_cast<const float4*>(P + 4 * g)[0];
        const float4 q23 = reinterpret_cast<const float4*>(P + 4 * g)[1];
        const float v0x = q01.x - ax, v0y = q01.y - ay, v1x = q01.z - ax, v1y = q01.w - ay;
        const float v2x = q23.x - ax, v2y = q23.y - ay, v3x = q23.z - ax, v3y = q23.w - ay;
        float d0 = ya_cross(vpx, vpy, v0x, v0y);
        const float d1 = ya_cross(v0x, v0y, v1x, v1y);
        const float d2 = ya_cross(v1x, v1y, v2x, v2y);
        const float d3 = ya_cross(v2x, v2y, v3x, v3y);
        if (g == 0 && d0 == 0.f) d0 = d1;   // point 359 repeats point 0 in the reference's wire format
        const uint32_t b0 = ya_bits(d0), b1 = ya_bits(d1), b2 = ya_bits(d2), b3 = ya_bits(d3), bp = ya_bits(pdb);
        const bool any_neg = ((b0 | b1 | b2 | b3 | bp) >> 31) != 0;
        const bool all_neg = ((b0 & b1 & b2 & b3 & bp) >> 31) != 0;
        // A group is redone point by point (phase C) when the direction changes inside it, or when it turns by
        // 180 degrees or more (then the sector test of its last point alone would miss rays).  The turn is only
        // looked at when the last point is closer to the anchor than 5 contour steps: otherwise every point of the
        // group and the previous one are at least 1 / 2 steps away and the four steps turn by less than
        // 60 + 3 * 29 degrees.  Same-direction steps vp->v1, v1->v3 (< 360 each) and vp->v3 all below 180 degrees
        // <=> the three cross products carry the direction's sign.
        bool rev = any_neg && !all_neg;
        if (!rev && fmaf(v3x, v3x, v3y * v3y) < near2) {
            const uint32_t c01 = ya_bits(ya_cross(vpx, vpy, v1x, v1y)), c13 = ya_bits(ya_cross(v1x, v1y, v3x, v3y));
            const uint32_t c03 = ya_bits(ya_cross(vpx, vpy, v3x, v3y));
            rev = any_neg ? (((c01 & c13 & c03) >> 31) == 0) : (((c01 | c13 | c03) >> 31) != 0);
        }
        if (rev) {
            if (nrev < YA_RMAX) sm.rev[nrev][tid] = (unsigned short)(g | (s << 7) | ((bp >> 31) << 14));
            else bad = true;
            ++nrev;
        }
        // sector of the group's last point; crossings are recorded unless phase C will redo the group
        const bool wup = rev ? (ya_cross(vpx, vpy, v3x, v3y) >= 0.f) : !any_neg;
        for (int guard = 0;; ++guard) {
            const int n = wup ? s + 1 : s;
            const float2 u = sm.raydir[n];
            const float x3 = ya_cross(u.x, u.y, v3x, v3y);
            if (!(wup ? (x3 >= 0.f) : (x3 < 0.f))) break;
            if (guard >= R) { bad = bad || !rev; break; }
            if (!rev) {
                const float x0 = ya_cross(u.x, u.y, v0x, v0y), x1 = ya_cross(u.x, u.y, v1x, v1y);
                const float x2 = ya_cross(u.x, u.y, v2x, v2y);
                const int k = wup ? ((x0 < 0.f) + (x1 < 0.f) + (x2 < 0.f)) : ((x0 >= 0.f) + (x1 >= 0.f) + (x2 >= 0.f));
                arc_log<R, NT>(sm, tid, (n == R) ? 0 : n, ya_wrap(4 * g - 1 + k), 0u);
            }
            s += wup ? 1 : -1;
            s = (s == R) ? 0 : ((s < 0) ? R - 1 : s);
        }
        if (rev) {   // several direction changes inside one group: make sure the sector is right
            const float2 lo = sm.raydir[s], hi = sm.raydir[s + 1];
            if (ya_cross(lo.x, lo.y, v3x, v3y) < 0.f || ya_cross(hi.x, hi.y, v3x, v3y) >= 0.f)
                s = arc_sector_of<R, NT>(sm, v3x, v3y);
        }
        pdb = d3;
        vpx = v3x;
        vpy = v3y;
    }
    sm.nrev[tid] = (unsigned char)((nrev < 255) ? nrev : 255);
    sm.nrpt[tid] = 0;
    sm.bad[tid] = bad ? 1 : 0;
    sm.anchor[tid] = make_float2(ax, ay);
}

// ------------------------------------------------------------------------------------------------
// phase C: one noted group, point by point
// ------------------------------------------------------------------------------------------------
template <int R, int NT>
YA_HD void arc_phase_c(ArcSmem<R, NT>& sm, int tid, int k, float ax, float ay) {
    const unsigned e = sm.rev[k][tid];
    const int g = e & 127u;
    int s = (e >> 7) & 127u;
    int pd = ((e >> 14) & 1u) ? -1 : 1;
    const float2* P = sm.cpad + YA_PAD;
    float vpx = P[4 * g - 1].x - ax, vpy = P[4 * g - 1].y - ay;
    int nrpt = sm.nrpt[tid];
    bool bad = false;
    for (int t = 0; t < 4; ++t) {
        const int q = ya_wrap(4 * g + t - 1);   // the step q -> q+1
        const float vx = P[4 * g + t].x - ax, vy = P[4 * g + t].y - ay;
        float d = ya_cross(vpx, vpy, vx, vy);
        if (g == 0 && t == 0 && d == 0.f) {      // same rule as the sweep
            const float nx = P[1].x - ax, ny = P[1].y - ay;
            d = ya_cross(vx, vy, nx, ny);
        }
        const int dirn = ((ya_bits(d) >> 31) != 0) ? -1 : 1;   // sign bit, as the sweep reads it
        if (dirn != pd) {
            // reversal at point q (sector s): a maximum leaves ray s+1 unreached, a minimum ray s
            const int ray = (pd > 0) ? ((s + 1 == R) ? 0 : s + 1) : s;
            arc_log<R, NT>(sm, tid, ray, q, YA_PSEUDO);
            if (nrpt < YA_RMAX) sm.rpt[nrpt][tid] = (unsigned short)q;
            else bad = true;
            ++nrpt;
            pd = dirn;
        }
        for (int guard = 0;; ++guard) {
            const int n = (dirn > 0) ? s + 1 : s;
            const float2 u = sm.raydir[n];
            const float x = ya_cross(u.x, u.y, vx, vy);
            if (!((dirn > 0) ? (x >= 0.f) : (x < 0.f))) break;
            if (guard >= R) { bad = true; break; }
            arc_log<R, NT>(sm, tid, (n == R) ? 0 : n, q, 0u);
            s += dirn;
            s = (s == R) ? 0 : ((s < 0) ? R - 1 : s);
        }
        vpx = vx;
        vpy = vy;
    }
    sm.nrpt[tid] = (unsigned char)((nrpt < 255) ? nrpt : 255);
    if (bad) sm.bad[tid] = 1;
}

// ------------------------------------------------------------------------------------------------
// pass 1: one crossing, six points
// ------------------------------------------------------------------------------------------------
// Returns true when the ray is settled (target written to sm.tv(i, tid)); false: the caller queues the ray.
template <int R, int NT>
YA_HD bool arc_pass1(ArcSmem<R, NT>& sm, const ArcConst& ac, int tid, int i, float ax, float ay, bool has_rev) {
    const int n = sm.cnt[i][tid];
    if (n == 0) {   // no arc crosses or ends next to this ray: every point is more than a ray spacing away
        sm.tv(i, tid) = YA_FLOOR;
        return true;
    }
    const unsigned ev = sm.slot[i][tid].x;
    const bool fast = (n == 1) && !(ev & YA_PSEUDO);
    const int c = ev & 511u;
    const float2 u = sm.raydir[i];
    const float2* W = sm.cpad + YA_PAD + c - 2;   // points c-2 .. c+3
    float key[6], d2[6], dot[6];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int t = 0; t < 6; ++t) {
        const float2 p = W[t];
        const float vx = p.x - ax, vy = p.y - ay;
        const float crs = ya_cross(u.x, u.y, vx, vy);
        dot[t] = fmaf(vx, u.x, vy * u.y);
        d2[t] = fmaf(vx, vx, vy * vy);
        key[t] = (crs * crs) * ya_rcp(d2[t]);
    }
    const float m = fmaxf(fmaxf(key[1], key[2]), fmaxf(key[3], key[4]));
    const float kmin = fminf(fminf(key[1], key[2]), fminf(key[3], key[4]));
    const float dmin = fminf(fminf(dot[1], dot[2]), fminf(dot[3], dot[4]));
    const float dd = fminf(fminf(d2[1], d2[2]), fminf(d2[3], d2[4]));
    const float far2 = fmaxf(fmaxf(d2[1], d2[2]), fmaxf(d2[3], d2[4]));
    // the comparisons are false for NaN keys (a contour point on the anchor), which sends the ray to pass 2
    const bool ok = fast && (key[0] >= m) && (key[5] >= m) && (dmin > 0.f) && (dd > 0.f) && (!has_rev || m <= ac.kstep) &&
                    (m == m);
    if (!ok) return false;
    sm.tv(i, tid) = (kmin > ac.g2) ? YA_FLOOR : fmaxf(sqrtf(far2), YA_FLOOR);
    return true;
}

// ------------------------------------------------------------------------------------------------
// pass 2: generic settlement of one (owner thread, ray) pair
// ------------------------------------------------------------------------------------------------
struct ArcTop4 {
    float k[4];
    int j[4];
};

YA_HD void top4_init(ArcTop4& T) {
    for (int i = 0; i < 4; ++i) { T.k[i] = 3.0f; T.j[i] = -1; }
}

// insert (key, point) unless the point is already listed; equal keys keep their order of arrival
YA_HD void top4_insert(ArcTop4& T, float key, int j) {
    if (j == T.j[0] || j == T.j[1] || j == T.j[2] || j == T.j[3]) return;
    if (!(key < T.k[3])) return;
    int pos = 3;
    while (pos > 0 && key < T.k[pos - 1]) { T.k[pos] = T.k[pos - 1]; T.j[pos] = T.j[pos - 1]; --pos; }
    T.k[pos] = key;
    T.j[pos] = j;
}

// key monotone in the angular distance over [0, 180] deg: sin^2 in front, 2 - sin^2 behind.  A contour point on
// the anchor itself has atan2(0, 0) = 0 in the reference: it is seen in the direction of ray 0.
YA_HD float arc_mono_key(float vx, float vy, const float2 u) {
    if (vx == 0.f && vy == 0.f) vx = 1.f;
    const float crs = ya_cross(u.x, u.y, vx, vy);
    const float dot = fmaf(vx, u.x, vy * u.y);
    const float d2 = fmaf(vx, vx, vy * vy);
    const float k = (crs * crs) / d2;
    return (dot > 0.f) ? k : 2.0f - k;
}

template <int R, int NT>
YA_HD float arc_top4_result(const ArcSmem<R, NT>& sm, const ArcConst& ac, const ArcTop4& T, float ax, float ay) {
    if (T.k[0] > ac.g2) return YA_FLOOR;
    float far2 = 0.f;
    for (int i = 0; i < 4; ++i) {
        if (T.j[i] < 0) continue;
        const float2 p = sm.cpad[YA_PAD + T.j[i]];
        const float vx = p.x - ax, vy = p.y - ay;
        far2 = fmaxf(far2, fmaf(vx, vx, vy * vy));
    }
    return fmaxf(sqrtf(far2), YA_FLOOR);
}

// evaluate contour indices lo..hi (relative to nothing: plain, may run outside [0,360)) into T and grow either
// end while its border point is nearer than the 4th key.  Returns false when the growth limit is hit.
template <int R, int NT>
YA_HD bool arc_window(const ArcSmem<R, NT>& sm, ArcTop4& T, const float2 u, float ax, float ay, int lo, int hi) {
    float klo = 0.f, khi = 0.f;
    for (int j = lo; j <= hi; ++j) {
        const int jj = ya_wrap(ya_wrap(j));
        const float2 p = sm.cpad[YA_PAD + jj];
        const float k = arc_mono_key(p.x - ax, p.y - ay, u);
        top4_insert(T, k, jj);
        if (j == lo) klo = k;
        khi = k;
    }
    int grown = 0;
    while (klo < T.k[3]) {
        if (++grown > YA_GROW) return false;
        const int jj = ya_wrap(ya_wrap(--lo));
        const float2 p = sm.cpad[YA_PAD + jj];
        klo = arc_mono_key(p.x - ax, p.y - ay, u);
        top4_insert(T, klo, jj);
    }
    grown = 0;
    while (khi < T.k[3]) {
        if (++grown > YA_GROW) return false;
        const int jj = ya_wrap(ya_wrap(++hi));
        const float2 p = sm.cpad[YA_PAD + jj];
        khi = arc_mono_key(p.x - ax, p.y - ay, u);
        top4_insert(T, khi, jj);
    }
    return true;
}

// exact scan of all points by one thread
template <int R, int NT>
YA_HD float arc_scan_serial(const ArcSmem<R, NT>& sm, const ArcConst& ac, int owner, int ray) {
    const float2 a = sm.anchor[owner];
    const float2 u = sm.raydir[ray];
    ArcTop4 T;
    top4_init(T);
    for (int j = 0; j < YCR_C; ++j) {
        const float2 p = sm.cpad[YA_PAD + j];
        const float k = arc_mono_key(p.x - a.x, p.y - a.y, u);
        if (k < T.k[3]) top4_insert(T, k, j);
    }
    return arc_top4_result<R, NT>(sm, ac, T, a.x, a.y);
}

// Returns false when the pair needs the exact scan.
template <int R, int NT>
YA_HD bool arc_settle_pair(const ArcSmem<R, NT>& sm, const ArcConst& ac, int owner, int ray, float& result) {
    const int n = sm.cnt[ray][owner];
    if (sm.bad[owner] || n > YA_KMAX || n == 0) return false;
    const float2 a = sm.anchor[owner];
    const float2 u = sm.raydir[ray];
    const ushort4 S = sm.slot[ray][owner];
    const unsigned ev[4] = {S.x, S.y, S.z, S.w};
    ArcTop4 T;
    top4_init(T);
    // A border certified against the 4th key of the moment stays certified: the key only falls as windows merge.
    for (int k = 0; k < n; ++k) {
        const int c = ev[k] & 511u;
        const bool pseudo = (ev[k] & YA_PSEUDO) != 0;
        if (!arc_window<R, NT>(sm, T, u, a.x, a.y, pseudo ? c - 3 : c - 2, c + 3)) return false;
    }
    if (sm.nrev[owner] > 0 && !(T.k[3] <= ac.kstep)) {
        // sparse ray: a reversal point whose pseudo-crossing went to another ray may still be among the four
        // nearest, so take the neighbourhoods of all of them
        const int np = sm.nrpt[owner];
        if (np > YA_RMAX) return false;
        for (int k = 0; k < np; ++k) {
            const int r = sm.rpt[k][owner];
            if (!arc_window<R, NT>(sm, T, u, a.x, a.y, r - 3, r + 3)) return false;
        }
    }
    result = arc_top4_result<R, NT>(sm, ac, T, a.x, a.y);
    return true;
}

// ------------------------------------------------------------------------------------------------
// per-candidate driver (host harness and device slow callers): everything for one thread, serially
// ------------------------------------------------------------------------------------------------
struct ArcStats { long long cand, rays_fast, rays_empty, rays_pair, rays_scan, nrev, bad; };

template <int R, int NT>
YA_HD void arc_candidate_serial(ArcSmem<R, NT>& sm, const ArcConst& ac, int tid, float ax, float ay, float near2,
                                float* t_out, ArcStats* st) {
    arc_sweep<R, NT>(sm, tid, ax, ay, near2);
    const int nrev = sm.nrev[tid];
    const int nloop = (nrev < YA_RMAX) ? nrev : YA_RMAX;
    for (int k = 0; k < nloop; ++k) arc_phase_c<R, NT>(sm, tid, k, ax, ay);
    if (st) { st->cand += 1; st->nrev += nrev; st->bad += sm.bad[tid]; }
    const bool bad = sm.bad[tid] != 0;
    for (int i = 0; i < R; ++i) {
        const int n = sm.cnt[i][tid];
        if (!bad && arc_pass1<R, NT>(sm, ac, tid, i, ax, ay, nrev > 0)) {
            if (st) { if (n == 0) st->rays_empty += 1; else st->rays_fast += 1; }
            t_out[i] = sm.tv(i, tid);
            continue;
        }
        float r;
        if (arc_settle_pair<R, NT>(sm, ac, tid, i, r)) {
            if (st) st->rays_pair += 1;
        } else {
            r = arc_scan_serial<R, NT>(sm, ac, tid, i);
            if (st) st->rays_scan += 1;
        }
        t_out[i] = r;
    }
}

template <int R, int NT>
YA_HD void arc_init_raydir_entry(ArcSmem<R, NT>& sm, int i) {
    const double ang = (double)((i % R) * (360 / R)) * (3.14159265358979323846 / 180.0);
    sm.raydir[i] = make_float2((float)cos(ang), (float)sin(ang));
}

// ------------------------------------------------------------------------------------------------
// device only: one warp = one chunk of 32 candidates of the same GT
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#ifndef YA_STATS
#define YA_STATS 0   // 1: count queued pairs / exact scans into g_ycr_stats (measurement builds only)
#endif

// Exact scan of all points for one pair by the whole warp (lanes take points lane, lane + 32, ...): each lane
// keeps its four smallest (key, index) words, then the four smallest of the warp are popped one by one.
template <int R, int NT>
__device__ __forceinline__ float arc_scan_warp(const ArcSmem<R, NT>& sm, const ArcConst& ac, int owner, int ray, unsigned lane) {
    const float2 a = sm.anchor[owner];
    const float2 u = sm.raydir[ray];
    const unsigned long long E = ~0ull;
    unsigned long long k0 = E, k1 = E, k2 = E, k3 = E;
    for (int j = (int)lane; j < YCR_C; j += 32) {
        const float2 p = sm.cpad[YA_PAD + j];
        const float k = arc_mono_key(p.x - a.x, p.y - a.y, u);
        const unsigned long long x = ((unsigned long long)__float_as_uint(k) << 32) | (unsigned)j;
        const unsigned long long n3 = max(k2, min(k3, x)), n2 = max(k1, min(k2, x)), n1 = max(k0, min(k1, x));
        k0 = min(k0, x); k1 = n1; k2 = n2; k3 = n3;
    }
    float far2 = 0.f, key0 = 3.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        unsigned long long mn = k0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (k0 == mn && mn != E) { k0 = k1; k1 = k2; k2 = k3; k3 = E; }
        if (mn != E) {
            if (r == 0) key0 = __uint_as_float((unsigned)(mn >> 32));
            const float2 p = sm.cpad[YA_PAD + (int)(mn & 0xffffffffu)];
            const float vx = p.x - a.x, vy = p.y - a.y;
            far2 = fmaxf(far2, fmaf(vx, vx, vy * vy));
        }
    }
    return (key0 > ac.g2) ? YA_FLOOR : fmaxf(sqrtf(far2), YA_FLOOR);
}

// Writes the contour of the chunk's GT (720 floats, x/y interleaved; lane l holds floats l, l+32, ... in reg[])
// into the padded copy and returns (warp-uniform) the square of the longest contour step in *l2max and whether
// some point may coincide with an anchor centre (both coordinates multiples of half the finest stride).
template <int R, int NT, int CW>
__device__ __forceinline__ bool arc_stage_contour(ArcSmem<R, NT>& sm, const float (&reg)[CW], unsigned lane, float inv_half_stride,
                                                  float* l2max) {
    float* dst = reinterpret_cast<float*>(sm.cpad + YA_PAD);
#pragma unroll
    for (int k = 0; k < CW; ++k)
        if (k * 32 + (int)lane < 2 * YCR_C) dst[k * 32 + lane] = reg[k];
    __syncwarp();
    if (lane < YA_PAD) {
        sm.cpad[lane] = sm.cpad[YCR_C + lane];                            // points 356..359 in front
        sm.cpad[YA_PAD + YCR_C + lane] = sm.cpad[YA_PAD + lane];          // points 0..3 behind
    }
    __syncwarp();
    float l2 = 0.f;
    bool lat = false;
    for (int j = (int)lane; j < YCR_C; j += 32) {
        const float2 p = sm.cpad[YA_PAD + j], q = sm.cpad[YA_PAD + j + 1];
        const float ex = q.x - p.x, ey = q.y - p.y;
        l2 = fmaxf(l2, fmaf(ex, ex, ey * ey));
        const float fx = p.x * inv_half_stride, fy = p.y * inv_half_stride;
        lat = lat || (fx == floorf(fx) && fy == floorf(fy));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l2 = fmaxf(l2, __shfl_xor_sync(0xffffffffu, l2, o));
    *l2max = l2;
    return __any_sync(0xffffffffu, lat);
}

// Ray targets of the warp's 32 candidates: on return sm.tv(i, lane) holds the target of ray i for every active
// lane.  `lattice`: see arc_stage_contour.  Returns the number of queued pairs (statistics).
template <int R, int NT>
__device__ __forceinline__ int arc_chunk_targets(ArcSmem<R, NT>& sm, const ArcConst& ac, int tid, bool active, float ax, float ay,
                                                 float near2, bool lattice, int* nscan_out) {
    const unsigned lane = tid & 31u;
    unsigned short* wq = sm.queue[tid >> 5];
    int nrev = 0;
    if (active) {
        arc_sweep<R, NT>(sm, tid, ax, ay, near2);
        nrev = min((int)sm.nrev[tid], YA_RMAX);
        if (lattice) {   // a contour point on the anchor has angle 0 in the reference whatever arc it lies on
            for (int j = 0; j < YCR_C; ++j) {
                const float2 p = sm.cpad[YA_PAD + j];
                if (p.x == ax && p.y == ay) sm.bad[tid] = 1;
            }
        }
    }
    const int nmax = __reduce_max_sync(0xffffffffu, nrev);
    for (int k = 0; k < nmax; ++k)
        if (k < nrev) arc_phase_c<R, NT>(sm, tid, k, ax, ay);
    const bool good = active && (sm.bad[tid] == 0);
    const bool has_rev = nrev > 0;
    int nq = 0;
#pragma unroll 2
    for (int i = 0; i < R; ++i) {
        bool unsettled = false;
        if (active) unsettled = !(good && arc_pass1<R, NT>(sm, ac, tid, i, ax, ay, has_rev));
        const unsigned ball = __ballot_sync(0xffffffffu, unsettled);
        if (ball) {
            if (unsettled) wq[nq + __popc(ball & ((1u << lane) - 1u))] = (unsigned short)((tid << 7) | i);
            nq += __popc(ball);
        }
    }
    __syncwarp();
    int nscan = 0;
    for (int q0 = 0; q0 < nq; q0 += 32) {
        const int q = q0 + (int)lane;
        unsigned e = 0;
        bool failed = false;
        if (q < nq) {
            e = wq[q];
            float t;
            if (arc_settle_pair<R, NT>(sm, ac, (int)(e >> 7), (int)(e & 127u), t)) sm.tv(e & 127u, e >> 7) = t;
            else failed = true;
        }
        unsigned fm = __ballot_sync(0xffffffffu, failed);
        nscan += __popc(fm);
        while (fm) {
            const int src = __ffs(fm) - 1;
            fm &= fm - 1;
            const unsigned es = __shfl_sync(0xffffffffu, e, src);
            const float t = arc_scan_warp<R, NT>(sm, ac, (int)(es >> 7), (int)(es & 127u), lane);
            if (lane == 0) sm.tv(es & 127u, es >> 7) = t;
        }
    }
    __syncwarp();
    if (nscan_out) *nscan_out = nscan;
    return nq;
}
#endif  // __CUDACC__
