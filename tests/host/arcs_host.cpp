// Host build of the per-thread logic of csrc/polar_arcs.cuh (test infrastructure): the same source the CUDA
// kernel compiles, run serially on the CPU so that tests/test_arcs_host.py can check it against the oracle
// without a GPU.  g++ -O2 -shared -fPIC -I/usr/local/cuda/include tests/host/arcs_host.cpp
#include "../../yolo-contour-regression_b200/csrc/polar_arcs.cuh"
#include <string.h>

template <int R>
static void run(const float* anchors, const float* contours, int M, float* t_out, long long* stats) {
    static ArcSmem<R, 1> sm;
    for (int i = 0; i <= R; ++i) arc_init_raydir_entry<R, 1>(sm, i);
    const ArcConst ac = make_arc_const(R);
    ArcStats st;
    memset(&st, 0, sizeof(st));
    for (int m = 0; m < M; ++m) {
        const float* c = contours + (size_t)m * 2 * YCR_C;
        for (int k = 0; k < YCR_C + 2 * YA_PAD; ++k) {
            const int j = ((k - YA_PAD) % YCR_C + YCR_C) % YCR_C;
            sm.cpad[k] = make_float2(c[2 * j], c[2 * j + 1]);
        }
        float l2 = 0.f;   // longest contour step, squared
        for (int j = 0; j < YCR_C; ++j) {
            const float ex = c[2 * ((j + 1) % YCR_C)] - c[2 * j], ey = c[2 * ((j + 1) % YCR_C) + 1] - c[2 * j + 1];
            l2 = fmaxf(l2, ex * ex + ey * ey);
        }
        arc_candidate_serial<R, 1>(sm, ac, 0, anchors[2 * m], anchors[2 * m + 1], 25.f * l2, t_out + (size_t)m * R, &st);
    }
    if (stats) {
        stats[0] = st.cand; stats[1] = st.rays_fast; stats[2] = st.rays_empty; stats[3] = st.rays_pair;
        stats[4] = st.rays_scan; stats[5] = st.nrev; stats[6] = st.bad;
    }
}

extern "C" int arcs_polar_targets(const float* anchors, const float* contours, int M, int R, float* t_out,
                                  long long* stats) {
    if (R == 36) run<36>(anchors, contours, M, t_out, stats);
    else if (R == 72) run<72>(anchors, contours, M, t_out, stats);
    else return -1;
    return 0;
}
