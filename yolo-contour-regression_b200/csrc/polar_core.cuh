// Polygon -> polar ray targets for one anchor per thread (reference semantics of
// utils/tal.py:1257-1277 == 1172-1193: for each of R fixed angles take the 4 contour points nearest
// in angle, target = max of their distances, 1e-6 when the nearest is more than 3 degrees away).
//
// The reference evaluates an (M,R,360) angle-difference tensor and top-k's it.  Here each thread
// sweeps the 360 contour points ONCE (they are broadcast from shared memory to the whole warp):
//   * a point belongs to the bin of its nearest ray (|delta| <= 180/R deg, decided by a cross/dot
//     test against the current ray direction - no atan2 anywhere);
//   * the bin's four nearest points live in registers as packed (23-bit fixed-point |sin delta|,
//     9-bit point index) keys and are kept sorted with 8 integer min/max per point; they are
//     spilled to / refilled from a per-thread shared-memory slot only when the sweep enters another
//     bin, and every bin change is logged as a (bin, first point) segment record;
//   * after the sweep a ray is settled when its own bin certifies the answer (four points strictly
//     inside the bin, or nothing within 3 degrees); the few others (sparse side of the contour) are
//     queued block-wide and settled exactly by widening the window bin by bin over the segment
//     records, with a pseudo-angle key that is monotone over [0,180] degrees.
// Every path is exact with respect to the reference whenever the reference's own selection is not
// within ~1e-5 degrees of a tie (the parity tests' margin checker uses 2e-4 degrees).
#pragma once
#include "common.cuh"

#define YCR_SEGCAP 64
#define YCR_NSCHED 5

struct PolarConst {
    float tan_in;       // tan(hw + 0.01 deg): bin membership |crs| <= tan_in * dot
    float key_scale;    // fixed-point scale of |sin delta| so that in-bin keys fit 23 bits
    uint32_t q_res;     // fixed(sin(hw - 0.02 deg)): 4th key below this => top-4 certified
    uint32_t q_gate;    // fixed(sin(3 deg))
    int gate_l1;        // 1 when hw - 0.01 > 3: an own bin without a point <= 3 deg certifies the gate
    int sched[YCR_NSCHED];      // window growth in bins
    float pk_win[YCR_NSCHED];   // pseudo-angle of ((2m+1)*hw - 0.02 deg)
    int gate_ok[YCR_NSCHED];    // (2m+1)*hw - 0.02 > 3
    float pk_gate;              // pseudo-angle of 3 deg
};

template <int R, int NT>
struct PolarSmem {
    uint4 list[R][NT];                 // per-thread, per-ray sorted packed keys (later: .x = target bits)
    float2 contour[YCR_C];
    float2 raydir[R];                  // (cos, sin) of i*360/R deg
    float2 anchor[NT];
    unsigned short seg[YCR_SEGCAP][NT];
    unsigned short nseg[NT];
    unsigned short queue[NT * R];      // (thread << 7) | ray
    int qcount;
    int work;                          // broadcast slot for the persistent loop
};

__device__ __forceinline__ void insert4(uint32_t& k0, uint32_t& k1, uint32_t& k2, uint32_t& k3, uint32_t x) {
    uint32_t lo;
    lo = min(k0, x); x = max(k0, x); k0 = lo;
    lo = min(k1, x); x = max(k1, x); k1 = lo;
    lo = min(k2, x); x = max(k2, x); k2 = lo;
    k3 = min(k3, x);
}

// Monotone map of the angle between v and the ray onto [0,4], from q=|cross| and d=dot.
__device__ __forceinline__ float pseudo_angle(float q, float d) {
    if (d >= q) return (d > 0.f) ? q / d : 0.f;
    if (d > -q) return 2.f - d / q;
    return 4.f + q / d;
}

// One sweep over the contour for the anchor (ax, ay) of this thread.
template <int R, int NT>
__device__ __forceinline__ void polar_sweep(PolarSmem<R, NT>& sm, const PolarConst& pc, int tid, float ax, float ay) {
#pragma unroll 4
    for (int i = 0; i < R; ++i) sm.list[i][tid] = make_uint4(YCR_EMPTY, YCR_EMPTY, YCR_EMPTY, YCR_EMPTY);
    int ray = 0;
    float cr = 1.f, sr = 0.f;
    uint32_t k0 = YCR_EMPTY, k1 = YCR_EMPTY, k2 = YCR_EMPTY, k3 = YCR_EMPTY;
    int nseg = 1;
    sm.seg[0][tid] = 0;  // (bin 0, first point 0)
    const float tan_in = pc.tan_in, ks = pc.key_scale;
#pragma unroll 4
    for (int j = 0; j < YCR_C; ++j) {
        const float2 p = sm.contour[j];
        float vx = p.x - ax, vy = p.y - ay;
        float l2 = fmaf(vx, vx, vy * vy);
        if (l2 == 0.f) { vx = 1.f; l2 = 1.f; }  // atan2(0,0) = 0: direction of ray 0
        const float inv = rsqrtf(l2);
        float dot = fmaf(vx, cr, vy * sr);
        float crs = fmaf(vy, cr, -vx * sr);
        if (!(fabsf(crs) <= tan_in * dot)) {
            sm.list[ray][tid] = make_uint4(k0, k1, k2, k3);
            int guard = 0;
            do {
                ray += (crs >= 0.f) ? 1 : -1;
                ray = (ray < 0) ? ray + R : ((ray >= R) ? ray - R : ray);
                const float2 cs = sm.raydir[ray];
                cr = cs.x; sr = cs.y;
                dot = fmaf(vx, cr, vy * sr);
                crs = fmaf(vy, cr, -vx * sr);
            } while (!(fabsf(crs) <= tan_in * dot) && ++guard < R);
            const uint4 L = sm.list[ray][tid];
            k0 = L.x; k1 = L.y; k2 = L.z; k3 = L.w;
            if (nseg < YCR_SEGCAP) sm.seg[nseg][tid] = (unsigned short)(ray | (j << 7));
            ++nseg;
        }
        const float key = fabsf(crs) * inv;
        const uint32_t pk = (__float_as_uint(fmaf(key, ks, 8388608.f)) << 9) | (uint32_t)j;
        insert4(k0, k1, k2, k3, pk);
    }
    sm.list[ray][tid] = make_uint4(k0, k1, k2, k3);
    sm.nseg[tid] = (unsigned short)min(nseg, 65535);
    sm.anchor[tid] = make_float2(ax, ay);
}

template <int R, int NT>
__device__ __forceinline__ float dist2_of(const PolarSmem<R, NT>& sm, uint32_t packed, float ax, float ay) {
    const float2 p = sm.contour[packed & 511u];
    const float vx = p.x - ax, vy = p.y - ay;
    return fmaf(vx, vx, vy * vy);
}

// Own-bin settlement of every ray of this thread; unsettled rays go to the block queue.
// On return sm.list[i][tid].x holds the float bits of the target for settled rays.
template <int R, int NT>
__device__ __forceinline__ void polar_settle_own(PolarSmem<R, NT>& sm, const PolarConst& pc, int tid, bool active,
                                                 float ax, float ay) {
    const unsigned lane = threadIdx.x & 31u;
    for (int i = 0; i < R; ++i) {
        bool unsettled = false;
        if (active) {
            const uint4 L = sm.list[i][tid];
            const bool own_gate = (L.x == YCR_EMPTY) || ((L.x >> 9) > pc.q_gate);
            if (own_gate && pc.gate_l1) {
                sm.list[i][tid].x = __float_as_uint(YCR_FLOOR);
            } else if (L.w != YCR_EMPTY && (L.w >> 9) < pc.q_res) {
                // (for hw < 3 deg every in-bin key is below the gate, so the gate cannot fire here)
                float m = dist2_of(sm, L.x, ax, ay);
                m = fmaxf(m, dist2_of(sm, L.y, ax, ay));
                m = fmaxf(m, dist2_of(sm, L.z, ax, ay));
                m = fmaxf(m, dist2_of(sm, L.w, ax, ay));
                sm.list[i][tid].x = __float_as_uint(fmaxf(sqrtf(m), YCR_FLOOR));
            } else {
                unsettled = true;
            }
        }
        const unsigned ball = __ballot_sync(0xffffffffu, unsettled);
        if (ball) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&sm.qcount, __popc(ball));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (unsettled) sm.queue[base + __popc(ball & ((1u << lane) - 1u))] = (unsigned short)((tid << 7) | i);
        }
    }
}

// Exact settlement of one queued (owner thread, ray) pair by window growth over segment records.
template <int R, int NT>
__device__ __noinline__ float polar_settle_pair(const PolarSmem<R, NT>& sm, const PolarConst& pc, int owner, int ray) {
    const float2 a = sm.anchor[owner];
    const float2 cs = sm.raydir[ray];
    float fk0 = 1e30f, fk1 = 1e30f, fk2 = 1e30f, fk3 = 1e30f;
    float fd0 = 0.f, fd1 = 0.f, fd2 = 0.f, fd3 = 0.f;  // squared distances ride along
    auto eval = [&](int j) {
        const float2 p = sm.contour[j];
        float vx = p.x - a.x, vy = p.y - a.y;
        const float l2 = fmaf(vx, vx, vy * vy);
        if (l2 == 0.f) vx = 1.f;
        const float d = fmaf(vx, cs.x, vy * cs.y);
        const float q = fabsf(fmaf(vy, cs.x, -vx * cs.y));
        float k = pseudo_angle(q, d), v = l2;
        // sorted insert of (k, v)
        if (k < fk3) {
            if (k < fk0) { float t = fk0; fk0 = k; k = t; t = fd0; fd0 = v; v = t; }
            if (k < fk1) { float t = fk1; fk1 = k; k = t; t = fd1; fd1 = v; v = t; }
            if (k < fk2) { float t = fk2; fk2 = k; k = t; t = fd2; fd2 = v; v = t; }
            if (k < fk3) { fk3 = k; fd3 = v; }
        }
    };
    const int nrec = sm.nseg[owner];
    if (nrec > YCR_SEGCAP) {
        for (int j = 0; j < YCR_C; ++j) eval(j);  // record overflow: plain exact scan
    } else {
        const uint4 L = sm.list[ray][owner];
        if (L.x != YCR_EMPTY) eval(L.x & 511u);
        if (L.y != YCR_EMPTY) eval(L.y & 511u);
        if (L.z != YCR_EMPTY) eval(L.z & 511u);
        if (L.w != YCR_EMPTY) eval(L.w & 511u);
        int mprev = 0;
        for (int s = 0; s < YCR_NSCHED; ++s) {
            const int m = pc.sched[s];
            for (int k = 0; k < nrec; ++k) {
                const unsigned rec = sm.seg[k][owner];
                int db = abs((int)(rec & 127u) - ray);
                db = min(db, R - db);
                if (db > mprev && db <= m) {
                    const int st = rec >> 7;
                    const int en = (k + 1 < nrec) ? (sm.seg[k + 1][owner] >> 7) : YCR_C;
                    for (int j = st; j < en; ++j) eval(j);
                }
            }
            mprev = m;
            if (2 * m + 1 >= R) break;                       // the whole circle is covered
            if (fk3 < pc.pk_win[s]) break;                   // four points certified inside the window
            if (pc.gate_ok[s] && fk0 > pc.pk_gate) break;    // nothing within 3 degrees, certified
        }
    }
    if (fk0 > pc.pk_gate) return YCR_FLOOR;
    const float m2 = fmaxf(fmaxf(fd0, fd1), fmaxf(fd2, fd3));
    return fmaxf(sqrtf(m2), YCR_FLOOR);
}

// Block-wide: settle all queued pairs densely (any thread may serve any owner).
template <int R, int NT>
__device__ __forceinline__ void polar_settle_queue(PolarSmem<R, NT>& sm, const PolarConst& pc, int tid) {
    const int nq = sm.qcount;
    for (int q = tid; q < nq; q += NT) {
        const unsigned e = sm.queue[q];
        const int owner = e >> 7, ray = e & 127u;
        const float t = polar_settle_pair<R, NT>(sm, pc, owner, ray);
        sm.list[ray][owner].x = __float_as_uint(t);
    }
}

static inline double ycr_deg2rad(double d) { return d * 3.14159265358979323846 / 180.0; }

static inline double ycr_pseudo_host(double deg) {
    const double q = sin(ycr_deg2rad(deg)), d = cos(ycr_deg2rad(deg));
    if (d >= q) return q / d;
    if (d > -q) return 2.0 - d / q;
    return 4.0 + q / d;
}

static inline PolarConst make_polar_const(int R) {
    PolarConst pc{};
    const double hw = 180.0 / R;
    pc.tan_in = (float)tan(ycr_deg2rad(hw + 0.01));
    const double scale = 8388607.0 / sin(ycr_deg2rad(hw + 0.02));
    pc.key_scale = (float)scale;
    pc.q_res = (uint32_t)(sin(ycr_deg2rad(hw - 0.02)) * scale);
    pc.q_gate = (uint32_t)(sin(ycr_deg2rad(YCR_GATE_DEG)) * scale);
    if (YCR_GATE_DEG >= hw + 0.01) pc.q_gate = 0x7FFFFFu;  // every in-bin key is below the gate
    pc.gate_l1 = (hw - 0.01 > YCR_GATE_DEG) ? 1 : 0;
    const int sched[YCR_NSCHED] = {1, 2, 4, 8, R / 2};
    for (int s = 0; s < YCR_NSCHED; ++s) {
        pc.sched[s] = sched[s];
        double win = (2 * sched[s] + 1) * hw - 0.02;
        if (win > 179.9) win = 179.9;
        pc.pk_win[s] = (float)ycr_pseudo_host(win);
        pc.gate_ok[s] = (win > YCR_GATE_DEG) ? 1 : 0;
    }
    pc.pk_gate = (float)ycr_pseudo_host(YCR_GATE_DEG);
    return pc;
}
