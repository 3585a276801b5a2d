// Polygon -> polar ray targets for one anchor per thread (reference semantics of
// utils/tal.py:1257-1277 == 1172-1193: for each of R fixed angles take the 4 contour points nearest
// in angle, target = max of their distances, 1e-6 when the nearest is more than 3 degrees away).
//
// The reference evaluates an (M,R,360) angle-difference tensor and top-k's it.  Here each thread
// sweeps the 360 contour points ONCE (they are broadcast from shared memory to the whole warp):
//   * a point belongs to the bin of its nearest ray (|delta| <= 180/R deg, decided by a cross/dot
//     test against the current ray direction - no atan2 anywhere);
//   * every bin keeps its four nearest points as packed (23-bit fixed-point |sin delta|, 9-bit point
//     index) keys in a per-thread uint4 slot of shared memory, kept sorted with 7 independent integer
//     min/max per point, and the exact number of points that fell into it (a byte; a bin that takes
//     256 points or more is detected through the total);
//   * after the sweep a ray is settled when its own bin certifies the answer (four points strictly
//     inside the bin, or nothing within 3 degrees).  The others (sparse side of the contour) go to the
//     queue of the thread's warp; a queued pair is settled by evaluating the contour neighbourhood of
//     the points its bin did catch, and the result is certified against the per-bin counts (every point
//     of the bins that could hold a nearer point has been looked at).  What cannot be certified is
//     settled right away by an exact scan of all 360 points, the warp's lanes taking 32 points at a time.
// Every path is exact with respect to the reference whenever the reference's own selection is not
// within ~1e-5 degrees of a tie (the parity tests' margin checker uses 2e-4 degrees).
#pragma once
#include "common.cuh"

// Angular slack of every bin / window certificate, in degrees: far above the fp32 noise of the cross/dot
// tests (~1e-5 deg).  It only decides WHICH exact path settles a ray, never the result; a point inside the
// +-3 TOL zone around a window edge sends the ray to the exact scan.
#define YCR_TOL_DEG 0.001
#ifndef YCR_GROUP
#define YCR_GROUP 4   // contour points fetched / inserted / stored together in the sweep
#endif
#ifndef YCR_FWD
#define YCR_FWD 2     // points of a group whose list updates are forwarded in registers
#endif
#define YCR_STR(x) #x
#define YCR_UNROLL(n) _Pragma(YCR_STR(unroll n))
#ifndef YCR_STREAMS
#define YCR_STREAMS 1   // independent ray trackers per thread (arcs of the contour swept in parallel; 1 measured best)
#endif
#ifndef YCR_OWN_UNROLL
#define YCR_OWN_UNROLL 2
#endif
#define YCR_MAXWIN 40  // window table entries (>= R/2 + 1 for R <= 72)
#ifndef YCR_NBR
#define YCR_NBR 4      // contour neighbours looked at on each side of a seed point (2: 1.006 ms, 3: 1.000, 4: 0.994, 5: 0.995)
#endif

struct PolarConst {
    float tan_in;       // tan(hw + TOL): bin membership |crs| <= tan_in * dot
    float cos_step, sin_step;  // cos/sin of the ray spacing 360/R deg
    float key_scale;    // fixed-point scale of |sin delta| so that in-bin keys fit 23 bits
    uint32_t q_res;     // fixed(sin(hw - 2 TOL)): 4th key below this => top-4 certified
    uint32_t q_gate;    // fixed(sin(3 deg))
    int gate_l1;        // 1 when hw - TOL > 3: an own bin without a point <= 3 deg certifies the gate
    int empty3_gate;    // 1 when 3*hw - 3 TOL > 3: three empty bins certify the gate
    int nwin;           // R/2 + 1 windows: window m covers bins ray-m .. ray+m
    float pk_lo[YCR_MAXWIN];  // pseudo-angle of ((2m+1)*hw - 3 TOL)
    float pk_hi[YCR_MAXWIN];  // pseudo-angle of ((2m+1)*hw + 3 TOL)
    float pk_gate;            // pseudo-angle of 3 deg
    float tan_win[YCR_MAXWIN];  // tan((2m+1)*hw + 5 TOL), 0 when that is 89 deg or more
    int m_gate;               // smallest window with (2m+1)*hw - 3 TOL > 3 deg
    uint32_t q2_gate;         // pseudo-angle of 3 deg in the 2^-21 fixed point of pack_pseudo
};

template <int R, int NT>
struct PolarSmem {
    uint4 list[R][NT];                 // per-thread, per-ray sorted packed keys
    float2 contour[YCR_C];
    float2 raydir[R];                  // (cos, sin) of i*360/R deg
    float2 anchor[NT];
    unsigned char cnt[R][NT];          // points per bin modulo 256 (255 = unknown, see polar_settle_own)
    unsigned short queue[NT / 32][16 * R];  // per warp: (thread << 7) | ray of the pairs the own bin could not settle
    // Settled ray target of (ray i, thread t).  It takes the place of the fourth key of that list: once a
    // ray is settled nothing reads its list again except the .x a neighbouring ray may borrow.
    __device__ __forceinline__ float& tv(int i, int t) { return reinterpret_cast<float*>(&list[i][t])[3]; }
    __device__ __forceinline__ float tv(int i, int t) const { return reinterpret_cast<const float*>(&list[i][t])[3]; }
};

// one MUFU.RSQ (the arguments here are squared pixel distances, never denormal)
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// sqrt of a squared pixel distance: one MUFU.SQRT (2 ulp; the ray targets are compared at 1e-5 and every discrete
// decision they feed is margin-checked at 2e-5) instead of the ~8 instructions of the correctly rounded sqrtf
#ifndef YCR_FAST_SQRT
#define YCR_FAST_SQRT 1
#endif
__device__ __forceinline__ float dist_sqrt(float x) {
#if YCR_FAST_SQRT
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}

// sorted insert with depth-2 dependency: new k_i = max(k_{i-1}, min(k_i, x))
__device__ __forceinline__ void insert4(uint32_t& k0, uint32_t& k1, uint32_t& k2, uint32_t& k3, uint32_t x) {
    const uint32_t n3 = max(k2, min(k3, x));
    const uint32_t n2 = max(k1, min(k2, x));
    const uint32_t n1 = max(k0, min(k1, x));
    k0 = min(k0, x);
    k1 = n1; k2 = n2; k3 = n3;
}

// Monotone map of the angle between v and the ray onto [0,4), from q=|cross| and d=dot: tan in the first
// octant pair, 2 - cot up to 135 deg, 4 + tan beyond.  Branch-free: one select chain, one fast division
// (2 ulp on a key whose decisive gaps are margin-checked at >= 1e-4 deg).
__device__ __forceinline__ float pseudo_angle(float q, float d) {
    const bool near = d >= q, mid = d > -q;
    const float num = near ? q : (mid ? -d : q);
    const float den = near ? d : (mid ? q : d);
    const float off = near ? 0.f : (mid ? 2.f : 4.f);
    const float k = off + ((den != 0.f) ? __fdividef(num, den) : 0.f);
    return fminf(k, 3.9999995f);
}

// pseudo-angle and point index as one sortable word: 23-bit fixed point (2^-21 units, ~3e-5 deg) | 9-bit index
__device__ __forceinline__ uint32_t pack_pseudo(float k, int j) {
    return (__float_as_uint(fmaf(k, 2097152.f, 8388608.f)) << 9) | (uint32_t)j;
}

// largest squared distance among the (up to four) contour points packed in K, seen from a
template <int R, int NT>
__device__ __forceinline__ float max_dist2(const PolarSmem<R, NT>& sm, const uint4 K, const float2 a) {
    const uint32_t e[4] = {K.x, K.y, K.z, K.w};
    float m = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (e[s] != YCR_EMPTY) {
            const float2 p = sm.contour[e[s] & 511u];
            const float vx = p.x - a.x, vy = p.y - a.y;
            m = fmaxf(m, fmaf(vx, vx, vy * vy));
        }
    }
    return m;
}

// One sweep over the contour for the anchor (ax, ay) of this thread.  The per-ray lists stay in
// shared memory and every point does a uniform read-insert-write on the list of its bin; the step
// to a neighbouring bin is branch-free, so the only divergent code is the walk over several bins
// (a lane whose anchor sits close to the contour).
// Bins narrower than the 3 degree gate (R > 60, e.g. 72 rays: +-2.5 deg) would rarely hold four points and
// never certify the gate.  There every point is inserted into the lists of BOTH rays that bracket it, so the
// list of ray i holds the nearest of all points within +-360/R deg (minus TOL) of it - the same situation as
// 36 rays with their +-5 deg bins; tracking, counts and windows keep using the disjoint half-spacing bins.
template <int R>
#ifndef YCR_DUAL_ALL
#define YCR_DUAL_ALL 0   // 1: also for wide bins (36 rays: queued pairs 4.1 -> 1.3 per candidate, kernel +1.6 %)
#endif
struct PolarDual { static constexpr bool value = YCR_DUAL_ALL || ((180.0 / R + YCR_TOL_DEG) <= YCR_GATE_DEG); };

template <int R, int NT>
__device__ __forceinline__ void polar_sweep(PolarSmem<R, NT>& sm, const PolarConst& pc, int tid, float ax, float ay) {
    constexpr bool kDual = PolarDual<R>::value;
#pragma unroll 4
    for (int i = 0; i < R; ++i) {
        sm.list[i][tid] = make_uint4(YCR_EMPTY, YCR_EMPTY, YCR_EMPTY, YCR_EMPTY);
        sm.cnt[i][tid] = 0;
    }
    // YCR_STREAMS independent trackers walk disjoint arcs of the contour (arc s starts at point s*C/S), so a
    // group holds S short serial chains instead of one long one; insertion order is irrelevant to the lists.
    constexpr int NS = YCR_STREAMS, PPS = YCR_GROUP / NS, SEG = YCR_C / NS;
    static_assert(YCR_GROUP % NS == 0 && YCR_C % NS == 0 && SEG % PPS == 0 && PPS % YCR_FWD == 0, "stream layout");
    int ray[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) ray[s] = 0;
    const float tan_in = pc.tan_in, ks = pc.key_scale, cstep = pc.cos_step, sstep = pc.sin_step;
    for (int j0 = 0; j0 < SEG; j0 += PPS) {
        // phase 1: everything that does not depend on the current bin, four points at once
        float vx[YCR_GROUP], vy[YCR_GROUP], inv[YCR_GROUP];
#pragma unroll
        for (int u = 0; u < YCR_GROUP; ++u) {
            const float2 p = sm.contour[(u / PPS) * SEG + j0 + (u % PPS)];
            // atan2(0,0) = 0, the direction of ray 0: a point exactly on the anchor gets vx = 1e-18 (any
            // other difference of coordinates is >= an ulp of a pixel coordinate and absorbs the bias)
            vx[u] = (p.x - ax) + 1e-18f;
            vy[u] = p.y - ay;
            inv[u] = rsqrt_fast(fmaf(vx[u], vx[u], vy[u] * vy[u]));
        }
        // phase 2: bin of each point (serial only through the tracked ray of its stream) and its key.
        // The usual move - one bin up or down - is branch-free: cross/dot and the ray direction are
        // rotated by one ray spacing in registers; the exact direction is reloaded from the table at
        // the start of every group, so at most PPS rotations (a few 1e-6 deg) ever accumulate.
        int rb[YCR_GROUP];
        uint32_t pk[YCR_GROUP];
        int rb2[YCR_GROUP];        // kDual: the other ray that brackets the point, and the key against it
        uint32_t pk2[YCR_GROUP];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            int ry = ray[s];
            const float2 cs0 = sm.raydir[ry];
            float cr = cs0.x, sr = cs0.y;
#pragma unroll
            for (int k = 0; k < PPS; ++k) {
                const int u = s * PPS + k;
                float dot = fmaf(vx[u], cr, vy[u] * sr);
                float crs = fmaf(vy[u], cr, -vx[u] * sr);
                const bool need = !(fabsf(crs) <= tan_in * dot);
                const bool up = crs >= 0.f;
                // rotate by zero (exact: x*1 + y*0) or by one ray spacing towards the point
                const float rc = need ? cstep : 1.f;
                const float rs = need ? (up ? sstep : -sstep) : 0.f;
                const float dot2 = fmaf(dot, rc, crs * rs);
                const float crs2 = fmaf(crs, rc, -dot * rs);
                const float cr2 = fmaf(cr, rc, -sr * rs);
                const float sr2 = fmaf(sr, rc, cr * rs);
                dot = dot2; crs = crs2; cr = cr2; sr = sr2;
                ry += need ? (up ? 1 : -1) : 0;
                ry = (ry < 0) ? R - 1 : ((ry >= R) ? 0 : ry);
                if (!(fabsf(crs) <= tan_in * dot)) {  // more than one bin away (sparse side / jump): walk, exactly
                    const int dir = up ? 1 : -1;
                    int guard = 0;
                    do {
                        ry += dir;
                        ry = (ry < 0) ? ry + R : ((ry >= R) ? ry - R : ry);
                        const float2 cs = sm.raydir[ry];
                        cr = cs.x; sr = cs.y;
                        dot = fmaf(vx[u], cr, vy[u] * sr);
                        crs = fmaf(vy[u], cr, -vx[u] * sr);
                    } while (!(fabsf(crs) <= tan_in * dot) && ++guard < R);
                }
                rb[u] = ry;
                const float key = fabsf(crs) * inv[u];
                pk[u] = (__float_as_uint(fmaf(key, ks, 8388608.f)) << 9) | (uint32_t)(s * SEG + j0 + k);
                if constexpr (kDual) {
                    const bool up2 = crs >= 0.f;
                    const float s2 = up2 ? sstep : -sstep;
                    const float crsn = fmaf(crs, cstep, -dot * s2);   // cross against the neighbouring ray
                    int r2 = ry + (up2 ? 1 : -1);
                    r2 = (r2 < 0) ? R - 1 : ((r2 >= R) ? 0 : r2);
                    rb2[u] = r2;
                    pk2[u] = (__float_as_uint(fmaf(fabsf(crsn) * inv[u], ks, 8388608.f)) << 9) | (uint32_t)(s * SEG + j0 + k);
                }
            }
            ray[s] = ry;
        }
        // phases 3-5, per sub-group of YCR_FWD points: the lists and counts are fetched together, updated in
        // contour order - forwarding the result of an earlier point of the same bin (the usual case) instead
        // of going through shared memory again - and written back in order (a later store to the same bin
        // carries all updates).  Sub-groups follow each other through shared memory.
#pragma unroll
        for (int u0 = 0; u0 < YCR_GROUP; u0 += YCR_FWD) {
            uint4 L[YCR_FWD];
            unsigned c[YCR_FWD];
#pragma unroll
            for (int u = 0; u < YCR_FWD; ++u) {
                L[u] = sm.list[rb[u0 + u]][tid];
                c[u] = sm.cnt[rb[u0 + u]][tid];
            }
#pragma unroll
            for (int u = 0; u < YCR_FWD; ++u) {
#pragma unroll
                for (int w = 0; w < u; ++w)
                    if (rb[u0 + w] == rb[u0 + u]) { L[u] = L[w]; c[u] = c[w]; }
                insert4(L[u].x, L[u].y, L[u].z, L[u].w, pk[u0 + u]);
                c[u] += 1u;   // stored modulo 256; polar_settle_own detects a wrapped bin by the total
            }
#pragma unroll
            for (int u = 0; u < YCR_FWD; ++u) {
                sm.list[rb[u0 + u]][tid] = L[u];
                sm.cnt[rb[u0 + u]][tid] = (unsigned char)c[u];
            }
        }
        if constexpr (kDual) {
            // second insertion (lists only - the counts stay those of the disjoint bins), one point after
            // the other through shared memory
#pragma unroll
            for (int u = 0; u < YCR_GROUP; ++u) {
                uint4 L = sm.list[rb2[u]][tid];
                insert4(L.x, L.y, L.z, L.w, pk2[u]);
                sm.list[rb2[u]][tid] = L;
            }
        }
    }
    sm.anchor[tid] = make_float2(ax, ay);
}


template <int R, int NT>
__device__ __forceinline__ float dist2_of(const PolarSmem<R, NT>& sm, uint32_t packed, float ax, float ay) {
    const float2 p = sm.contour[packed & 511u];
    const float vx = p.x - ax, vy = p.y - ay;
    return fmaf(vx, vx, vy * vy);
}

template <int R, int NT>
__device__ __forceinline__ bool polar_settle_pair(const PolarSmem<R, NT>& sm, const PolarConst& pc, int owner, int ray,
                                                  float& result);

// Exact scan of all points by one thread (only used when a warp queue overflows).
template <int R, int NT>
__device__ __noinline__ float polar_scan_serial(const PolarSmem<R, NT>& sm, const PolarConst& pc, int owner, int ray) {
    const float2 a = sm.anchor[owner];
    const float2 cs = sm.raydir[ray];
    uint4 K = make_uint4(YCR_EMPTY, YCR_EMPTY, YCR_EMPTY, YCR_EMPTY);
#pragma unroll 1
    for (int j = 0; j < YCR_C; ++j) {
        const float2 p = sm.contour[j];
        float vx = p.x - a.x;
        const float vy = p.y - a.y;
        if (vx == 0.f && vy == 0.f) vx = 1.f;
        insert4(K.x, K.y, K.z, K.w, pack_pseudo(pseudo_angle(fabsf(fmaf(vy, cs.x, -vx * cs.y)), fmaf(vx, cs.x, vy * cs.y)), j));
    }
    if ((K.x >> 9) > pc.q2_gate) return YCR_FLOOR;
    return fmaxf(dist_sqrt(max_dist2<R, NT>(sm, K, a)), YCR_FLOOR);
}

// Own-bin settlement of every ray of this thread; unsettled rays go to the queue of the thread's warp
// (warps never touch each other's candidates, so the whole settlement needs no block barrier).
// Returns the warp-uniform number of queued pairs.  On return sm.tv(i, tid) holds the target of every
// settled ray (the lists stay intact).
template <int R, int NT>
__device__ __forceinline__ int polar_settle_own(PolarSmem<R, NT>& sm, const PolarConst& pc, int tid, bool active,
                                                float ax, float ay) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned short* wq = sm.queue[tid >> 5];
    int nq = 0;
    unsigned csum = 0;
YCR_UNROLL(YCR_OWN_UNROLL)
    for (int i = 0; i < R; ++i) {
        bool unsettled = false;
        if (active) {
            const uint4 L = sm.list[i][tid];
            csum += sm.cnt[i][tid];
            const bool own_gate = (L.x == YCR_EMPTY) || ((L.x >> 9) > pc.q_gate);
            constexpr double kHwEff = PolarDual<R>::value ? 360.0 / R : 180.0 / R;     // reach of a list
            constexpr bool kGateL1 = (kHwEff - YCR_TOL_DEG) > YCR_GATE_DEG;             // == pc.gate_l1
            constexpr bool kEmpty3 = (3 * 180.0 / R - 3 * YCR_TOL_DEG) > YCR_GATE_DEG;  // == pc.empty3_gate
            bool gate = own_gate && kGateL1;
            if (!kGateL1 && kEmpty3 && L.x == YCR_EMPTY) {
                // bins narrower than the 3 degree gate: an empty own bin between two empty neighbours
                // certifies it (nothing within 3*hw - 3 TOL > 3 degrees of the ray)
                const int ip = (i == 0) ? R - 1 : i - 1, in = (i == R - 1) ? 0 : i + 1;
                gate = (sm.cnt[ip][tid] | sm.cnt[in][tid]) == 0;
            }
            if (gate) {
                sm.tv(i, tid) = YCR_FLOOR;
            } else if (L.w != YCR_EMPTY && (L.w >> 9) < pc.q_res) {
                // (for hw < 3 deg every in-bin key is below the gate, so the gate cannot fire here)
                float m = dist2_of(sm, L.x, ax, ay);
                m = fmaxf(m, dist2_of(sm, L.y, ax, ay));
                m = fmaxf(m, dist2_of(sm, L.z, ax, ay));
                m = fmaxf(m, dist2_of(sm, L.w, ax, ay));
                sm.tv(i, tid) = fmaxf(dist_sqrt(m), YCR_FLOOR);
            } else {
                unsettled = true;
            }
        }
        const unsigned ball = __ballot_sync(0xffffffffu, unsettled);
        if (ball) {
            if (unsettled) {
                const int slot = nq + __popc(ball & ((1u << lane) - 1u));
                if (slot < 16 * R) {
                    wq[slot] = (unsigned short)((tid << 7) | i);
                } else {  // queue full (pathological warp): settle right here, serially
                    sm.tv(i, tid) = polar_scan_serial<R, NT>(sm, pc, tid, i);
                }
            }
            nq += __popc(ball);
        }
    }
    if (active && csum != YCR_C) {
        // a bin took 256 points or more and its byte count wrapped (the counts must add up to the contour):
        // mark every bin saturated, which sends this thread's queued pairs to the exact scan
        for (int i = 0; i < R; ++i) sm.cnt[i][tid] = 255;
    }
    return min(nq, 16 * R);
}

// Smallest window (in bins on each side of the ray) whose bins hold at least four points and that
// reaches past the 3 degree gate; its point count.  Returns false when a count is saturated or the
// window would be the whole circle.
template <int R, int NT>
__device__ __forceinline__ bool polar_window(const PolarSmem<R, NT>& sm, const PolarConst& pc, int owner, int ray,
                                             int& m_out, int& n_out) {
    int nb = sm.cnt[ray][owner];
    bool sat = nb == 255;
    int m = 0;
    while ((nb < 4 || m < pc.m_gate) && 2 * (m + 1) < R) {
        ++m;
        const int c1 = sm.cnt[(ray + m) % R][owner], c2 = sm.cnt[(ray + R - m) % R][owner];
        sat |= (c1 == 255) | (c2 == 255);
        nb += c1 + c2;
    }
    m_out = m;
    n_out = nb;
    return !sat && nb >= 4;
}

// Neighbourhood settlement of one queued (owner thread, ray) pair.  Returns false when the result
// cannot be certified (the pair then takes the exact scan).
//   1. the per-bin counts give the window (bins ray-m .. ray+m) that must contain the four nearest points;
//   2. the points the own bin caught are normally consecutive contour indices: evaluate that index range
//      widened by YCR_NBR on both sides, then keep growing either end while its last point still lies
//      inside the window (along a locally monotone contour this visits exactly the points of the window);
//   3. certify: the number of evaluated points inside the window equals the number of points the sweep
//      counted in the window's bins, and none sits in the +-3 TOL fuzz zone of the window edge.
// Seeds spread over distant parts of the contour (several arcs through one bin) go to the exact scan.
#define YCR_GROW 20
template <int R, int NT>
__device__ __forceinline__ bool polar_settle_pair(const PolarSmem<R, NT>& sm, const PolarConst& pc, int owner, int ray,
                                                  float& result) {
    int m, nbins;
    if (!polar_window<R, NT>(sm, pc, owner, ray, m, nbins)) return false;
    const float win_lo = pc.pk_lo[m], win_hi = pc.pk_hi[m];
    const float2 a = sm.anchor[owner];
    const float2 cs = sm.raydir[ray];
    uint4 L = sm.list[ray][owner];
    if (L.x == YCR_EMPTY) {  // empty own bin (only reachable when hw < 3 deg): nearest of each neighbour bin
        L.x = sm.list[(ray + R - 1) % R][owner].x;
        L.y = sm.list[(ray + 1) % R][owner].x;
        if (L.x == YCR_EMPTY) { L.x = L.y; L.y = YCR_EMPTY; }
        if (L.x == YCR_EMPTY) return false;
    }
    // index range of the seeds on the circle of contour indices, measured from the first seed
    const int j0 = L.x & 511u;
    int dmin = 0, dmax = 0;
    {
        const uint32_t e[3] = {L.y, L.z, L.w};
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            if (e[s] != YCR_EMPTY) {
                int d = (int)(e[s] & 511u) - j0;
                d = (d > YCR_C / 2) ? d - YCR_C : ((d < -YCR_C / 2) ? d + YCR_C : d);
                dmin = min(dmin, d);
                dmax = max(dmax, d);
            }
        }
    }
    if (dmax - dmin > 16) return false;
    uint4 K = make_uint4(YCR_EMPTY, YCR_EMPTY, YCR_EMPTY, YCR_EMPTY);
    int n_in = 0, n_maybe = 0;
    auto eval = [&](int d) -> float {  // contour point j0 + d (mod C); branch-free
        int j = j0 + d;
        j = (j < 0) ? j + YCR_C : ((j >= YCR_C) ? j - YCR_C : j);
        const float2 p = sm.contour[j];
        float vx = p.x - a.x;
        const float vy = p.y - a.y;
        if (vx == 0.f && vy == 0.f) vx = 1.f;
        const float k = pseudo_angle(fabsf(fmaf(vy, cs.x, -vx * cs.y)), fmaf(vx, cs.x, vy * cs.y));
        n_in += (k < win_lo) ? 1 : 0;
        n_maybe += (k < win_hi) ? 1 : 0;
        insert4(K.x, K.y, K.z, K.w, pack_pseudo(k, j));
        return k;
    };
    int lo = dmin - YCR_NBR, hi = dmax + YCR_NBR;
    float klo = 0.f, khi = 0.f;
    for (int d = lo; d <= hi; ++d) {
        const float k = eval(d);
        if (d == lo) klo = k;
        khi = k;
    }
    int g = 0;
    while (klo < win_hi && g < YCR_GROW) { klo = eval(--lo); ++g; }
    int h = 0;
    while (khi < win_hi && h < YCR_GROW) { khi = eval(++hi); ++h; }
    if (g >= YCR_GROW || h >= YCR_GROW || n_in != n_maybe || n_in != nbins) return false;
    result = ((K.x >> 9) > pc.q2_gate) ? YCR_FLOOR : fmaxf(dist_sqrt(max_dist2<R, NT>(sm, K, a)), YCR_FLOOR);
    return true;
}

// Exact scan of all points for one pair, one warp per pair (lanes take consecutive points).  When the
// bin counts bound the window that holds the four nearest points, points outside it are skipped after
// a two-instruction test, so only the one or two iterations whose 32 consecutive points touch the
// window pay for key evaluation and insertion.  The four smallest of the lanes' packed keys are then
// popped with one warp min-reduction each (keys are unique: they carry the point index).
template <int R, int NT>
__device__ __forceinline__ float polar_scan_pair(const PolarSmem<R, NT>& sm, const PolarConst& pc, int owner, int ray,
                                                 unsigned lane) {
    const float2 a = sm.anchor[owner];
    const float2 cs = sm.raydir[ray];
    int m, nbins;
    float tanw = 0.f;
    if (polar_window<R, NT>(sm, pc, owner, ray, m, nbins)) tanw = pc.tan_win[m];  // 0: no usable bound
    uint4 K = make_uint4(YCR_EMPTY, YCR_EMPTY, YCR_EMPTY, YCR_EMPTY);
    for (int j = lane; j < YCR_C; j += 32) {
        const float2 p = sm.contour[j];
        float vx = p.x - a.x;
        const float vy = p.y - a.y;
        if (vx == 0.f && vy == 0.f) vx = 1.f;
        const float d = fmaf(vx, cs.x, vy * cs.y);
        const float q = fabsf(fmaf(vy, cs.x, -vx * cs.y));
        if (tanw == 0.f || q <= tanw * d) insert4(K.x, K.y, K.z, K.w, pack_pseudo(pseudo_angle(q, d), j));
    }
    uint4 W;
    uint32_t* w = &W.x;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t mn = __reduce_min_sync(0xffffffffu, K.x);
        if (K.x == mn && mn != YCR_EMPTY) { K.x = K.y; K.y = K.z; K.z = K.w; K.w = YCR_EMPTY; }
        w[r] = mn;
    }
    if ((W.x >> 9) > pc.q2_gate) return YCR_FLOOR;
    return fmaxf(dist_sqrt(max_dist2<R, NT>(sm, W, a)), YCR_FLOOR);
}


// Warp-wide: settle the warp's queued pairs, 32 at a time; a pair whose neighbourhood settlement cannot
// be certified is scanned exactly by the whole warp right away.  No block barrier.  Returns the number of
// exact scans (statistics).
template <int R, int NT>
__device__ __forceinline__ int polar_settle_batch(PolarSmem<R, NT>& sm, const PolarConst& pc, const unsigned short* wq,
                                                  int q0, int nq, unsigned lane) {
    const int q = q0 + (int)lane;
    unsigned e = 0;
    bool failed = false;
    if (q < nq) {
        e = wq[q];
        float t;
        if (polar_settle_pair<R, NT>(sm, pc, (int)(e >> 7), (int)(e & 127u), t)) sm.tv(e & 127u, e >> 7) = t;
        else failed = true;
    }
    unsigned fm = __ballot_sync(0xffffffffu, failed);
    const int nscan = __popc(fm);
    while (fm) {
        const int src = __ffs(fm) - 1;
        fm &= fm - 1;
        const unsigned es = __shfl_sync(0xffffffffu, e, src);
        const float t = polar_scan_pair<R, NT>(sm, pc, (int)(es >> 7), (int)(es & 127u), lane);
        if (lane == 0) sm.tv(es & 127u, es >> 7) = t;
    }
    return nscan;
}

template <int R, int NT>
__device__ __forceinline__ int polar_settle_queue(PolarSmem<R, NT>& sm, const PolarConst& pc, int tid, int nq) {
    const unsigned lane = tid & 31u;
    const unsigned short* wq = sm.queue[tid >> 5];
    int nscan = 0;
    __syncwarp();
    for (int q0 = 0; q0 < nq; q0 += 32) nscan += polar_settle_batch<R, NT>(sm, pc, wq, q0, nq, lane);
    __syncwarp();
    return nscan;
}

// Block-shared form for two-warp blocks: a warp that runs out of own pairs takes batches of the other
// warp's queue (queue lengths differ a lot between the two halves of a chunk).  ctl, per warp w:
// ctl[w] = published queue length (-1 until the warp's own-bin settlement is complete), ctl[2 + w] = next
// unclaimed entry, ctl[4 + w] = entries completed.  The caller resets ctl behind a block barrier.
// On return every pair of THIS warp's queue is settled, whoever did it.
template <int R, int NT>
__device__ __forceinline__ int polar_settle_queue_shared(PolarSmem<R, NT>& sm, const PolarConst& pc, int tid, int nq,
                                                         volatile int* ctl) {
    static_assert(NT == 64, "two warps");
    const unsigned lane = tid & 31u;
    const int w = tid >> 5;
    int nscan = 0;
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();   // lists, counts, anchors, queue entries before the published length
        ctl[w] = nq;
    }
    __syncwarp();
    for (int turn = 0; turn < 2; ++turn) {
        const int qw = w ^ turn;
        int n = nq;
        if (turn == 1) {
            // every lane polls (uniform control flow; the value changes once, from -1 to the length)
            // (back-off: an idle warp of a half-empty chunk waits here for the other warp's whole sweep, and
            // the kernel is issue-bound)
            for (unsigned ns = 64; (n = ctl[qw]) < 0; ns = min(ns * 2, 2048u)) __nanosleep(ns);
            __threadfence_block();
            __syncwarp();
        }
        const unsigned short* wq = sm.queue[qw];
        for (;;) {
            int q0 = 0;
            if (lane == 0) q0 = atomicAdd(const_cast<int*>(&ctl[2 + qw]), 32);
            q0 = __shfl_sync(0xffffffffu, q0, 0);
            if (q0 >= n) break;
            nscan += polar_settle_batch<R, NT>(sm, pc, wq, q0, n, lane);
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();   // results before the completion count
                atomicAdd(const_cast<int*>(&ctl[4 + qw]), min(32, n - q0));
            }
            __syncwarp();
        }
    }
    for (unsigned ns = 32; ctl[4 + w] < nq; ns = min(ns * 2, 1024u)) __nanosleep(ns);
    __threadfence_block();
    __syncwarp();
    return nscan;
}

static inline double ycr_deg2rad(double d) { return d * 3.14159265358979323846 / 180.0; }

static inline double ycr_pseudo_host(double deg) {
    if (deg > 179.99) deg = 179.99;
    const double q = sin(ycr_deg2rad(deg)), d = cos(ycr_deg2rad(deg));
    if (d >= q) return q / d;
    if (d > -q) return 2.0 - d / q;
    return 4.0 + q / d;
}

static inline PolarConst make_polar_const(int R) {
    PolarConst pc{};
    const double hw = 180.0 / R;
    const double T = YCR_TOL_DEG;
    pc.tan_in = (float)tan(ycr_deg2rad(hw + T));
    pc.cos_step = (float)cos(ycr_deg2rad(2 * hw));
    pc.sin_step = (float)sin(ycr_deg2rad(2 * hw));
    // reach of a list: the bin, or - bins narrower than the gate (PolarDual) - the two bins around the ray
    const double hw_eff = (YCR_DUAL_ALL || hw + T <= YCR_GATE_DEG) ? 2 * hw : hw;
    const double scale = 8388607.0 / sin(ycr_deg2rad(hw_eff + 2 * T));
    pc.key_scale = (float)scale;
    pc.q_res = (uint32_t)(sin(ycr_deg2rad(hw_eff - 2 * T)) * scale);
    pc.q_gate = (uint32_t)(sin(ycr_deg2rad(YCR_GATE_DEG)) * scale);
    if (YCR_GATE_DEG >= hw_eff + T) pc.q_gate = 0x7FFFFFu;  // every listed key is below the gate
    pc.gate_l1 = (hw_eff - T > YCR_GATE_DEG) ? 1 : 0;
    pc.empty3_gate = (3 * hw - 3 * T > YCR_GATE_DEG) ? 1 : 0;
    pc.nwin = R / 2 + 1;
    for (int m = 0; m < YCR_MAXWIN; ++m) {
        const double w = (2 * m + 1) * hw;
        pc.pk_lo[m] = (float)ycr_pseudo_host(w - 3 * T);
        pc.pk_hi[m] = (float)ycr_pseudo_host(w + 3 * T);
    }
    // the last window is the whole circle: nothing lies outside it
    pc.pk_lo[R / 2] = 5.f;
    pc.pk_hi[R / 2] = 5.f;
    pc.pk_gate = (float)ycr_pseudo_host(YCR_GATE_DEG);
    pc.q2_gate = (uint32_t)(ycr_pseudo_host(YCR_GATE_DEG) * 2097152.0);
    pc.m_gate = 0;
    while ((2 * pc.m_gate + 1) * hw - 3 * T <= YCR_GATE_DEG) ++pc.m_gate;
    for (int m = 0; m < YCR_MAXWIN; ++m) {
        const double w = (2 * m + 1) * hw + 5 * T;
        pc.tan_win[m] = (w < 89.0) ? (float)tan(ycr_deg2rad(w)) : 0.f;
    }
    return pc;
}
