/* Plain-C client of libce2e.so: proves the boundary needs no Python, no torch and no C++.
 *
 * Known answers are the hand-derivable ones of SURVEY.md section 8c (none is reference-pinned;
 * the pinned parity lives in tests/test_gpu_parity.py against the golden vectors):
 *   - action scaling (DM:128-132): [1,1] -> [0.4, 1.5]; [-1,-1] -> [-0.4, -3.0]; 2 clips to 1.05 -> 0.42
 *   - f_xu([5,0,0,0,0,90], [0,0], 0.1) (DM:52-83): straight ahead, y' = 0.5,
 *     x' = 0.1 * 5 * cos(fp32(pi/2)) = -2.1856e-8, lateral states stay 0
 *   - an ego on waypoint 10k of a straight path with the path's heading and v = 8 has a zero
 *     tracking error (DM:735-770), and find_closest_point returns 10k (DM:702-715)
 *   - a vehicle at the pad position (E2E:440-447) adds exactly 0 to veh2veh (DM:210-229)
 *   - error convention: NULL pointer -> CE2E_ERR_NULL, task 7 -> CE2E_ERR_TASK, path 9 -> CE2E_ERR_PATH
 *
 * Build (tests/test_gpu_parity.py::test_plain_c_client does this on the GPU box):
 *   gcc -std=c99 -Wall -Iinclude -I/usr/local/cuda/include tests/c_abi/kat.c \
 *       -Lenv_build_b200/csrc -lce2e -L/usr/local/cuda/lib64 -lcudart -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "ce2e.h"

static int failures = 0;

#define CHECK(cond, ...)                                  \
    do {                                                  \
        if (!(cond)) {                                    \
            ++failures;                                   \
            printf("FAIL %s:%d: ", __FILE__, __LINE__);   \
            printf(__VA_ARGS__);                          \
            printf("\n");                                 \
        }                                                 \
    } while (0)

#define CU(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); \
            exit(2);                                                               \
        }                                                                          \
    } while (0)

#define OK(call)                                                             \
    do {                                                                     \
        int rc_ = (call);                                                    \
        CHECK(rc_ == CE2E_OK, "%s -> %d (%s)", #call, rc_, ce2e_last_error()); \
    } while (0)

static float *to_dev(const float *h, size_t n) {
    float *d;
    CU(cudaMalloc((void **)&d, n * sizeof(float)));
    CU(cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice));
    return d;
}

static void to_host(float *h, const float *d, size_t n) {
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h, d, n * sizeof(float), cudaMemcpyDeviceToHost));
}

static void test_action_transform(void) {
    const float a[6] = {1.f, 1.f, -1.f, -1.f, 2.f, 0.f};
    float out[6];
    float *da = to_dev(a, 6), *dout = to_dev(a, 6);
    OK(ce2e_action_transform(da, dout, 3, NULL));
    to_host(out, dout, 6);
    CHECK(out[0] == 0.4f && out[1] == 2.25f * 1.f - 0.75f, "[1,1] -> %g %g", out[0], out[1]);
    CHECK(out[2] == -0.4f && out[3] == -3.0f, "[-1,-1] -> %g %g", out[2], out[3]);
    CHECK(out[4] == 0.4f * 1.05f && out[5] == -0.75f, "[2,0] -> %g %g", out[4], out[5]);
    cudaFree(da);
    cudaFree(dout);
}

static void test_dynamics(void) {
    const float s[6] = {5.f, 0.f, 0.f, 0.f, 0.f, 90.f}, a[2] = {0.f, 0.f};
    float nxt[6], par[4];
    float *ds = to_dev(s, 6), *da = to_dev(a, 2), *dn = to_dev(s, 6), *dp = to_dev(s, 4);
    OK(ce2e_dynamics_step(ds, 6, da, 0.1, dn, 6, dp, 0, 1, NULL));
    to_host(nxt, dn, 6);
    to_host(par, dp, 4);
    CHECK(nxt[0] == 5.f && nxt[1] == 0.f && nxt[2] == 0.f, "vx vy r = %g %g %g", nxt[0], nxt[1], nxt[2]);
    CHECK(fabsf(nxt[3] - (-2.1855694e-8f)) < 1e-12f, "x' = %.9g", nxt[3]);
    CHECK(nxt[4] == 0.5f && nxt[5] == 90.f, "y' phi' = %g %g", nxt[4], nxt[5]);
    CHECK(par[0] == 0.f && par[1] == 0.f, "slip angles %g %g", par[0], par[1]);
    /* ego_predict's clip (DM:390): a braking ego never rolls backwards */
    {
        const float s2[6] = {0.05f, 0.f, 0.f, 0.f, 0.f, 90.f}, a2[2] = {0.f, -3.f};
        CU(cudaMemcpy(ds, s2, sizeof(s2), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(da, a2, sizeof(a2), cudaMemcpyHostToDevice));
        OK(ce2e_dynamics_step(ds, 6, da, 0.1, dn, 6, NULL, 1, 1, NULL));
        to_host(nxt, dn, 6);
        CHECK(nxt[0] == 0.f, "clipped v_x = %g", nxt[0]);
    }
    cudaFree(ds); cudaFree(da); cudaFree(dn); cudaFree(dp);
}

/* a 3000-point straight path x = 1.875, y = -60 + k/30, heading 90 deg (like DM:604-611) */
#define NPT 3000
static ce2e_paths *make_paths(void) {
    static float xs[NPT], ys[NPT], ph[NPT];
    const float *px[1], *py[1], *pp[1];
    int32_t lens[1] = {NPT};
    ce2e_paths *h = NULL;
    int k;
    for (k = 0; k < NPT; ++k) {
        xs[k] = 1.875f;
        ys[k] = (float)(-60.0 + k / 30.0);
        ph[k] = 90.f;
    }
    px[0] = xs; py[0] = ys; pp[0] = ph;
    OK(ce2e_paths_create(CE2E_TASK_STRAIGHT, 1, lens, px, py, pp, &h));
    return h;
}

static void test_paths_and_rollout(void) {
    ce2e_paths *h = make_paths();
    const int k = 1230;                                  /* a multiple of 10 */
    const float wx = 1.875f, wy = (float)(-60.0 + k / 30.0);
    long long idx[2];
    float trk[3], pts[6];
    if (!h) return;
    {
        const float x[2] = {wx, wx + 0.3f}, y[2] = {wy, wy};
        float *dx = to_dev(x, 2), *dy = to_dev(y, 2), *dpts = to_dev(pts, 6);
        int64_t *didx;
        int brute;
        CU(cudaMalloc((void **)&didx, 2 * sizeof(int64_t)));
        for (brute = 0; brute < 2; ++brute) {
            OK(ce2e_find_closest_point(h, 0, dx, dy, 10, brute, didx, dpts, 2, NULL));
            CU(cudaDeviceSynchronize());
            CU(cudaMemcpy(idx, didx, sizeof(idx), cudaMemcpyDeviceToHost));
            to_host(pts, dpts, 6);
            CHECK(idx[0] == k && idx[1] == k, "closest index %lld %lld (brute=%d)", idx[0], idx[1], brute);
            CHECK(pts[0] == wx && pts[2] == wy && pts[4] == 90.f, "closest point %g %g %g", pts[0], pts[2], pts[4]);
        }
        cudaFree(dx); cudaFree(dy); cudaFree(dpts); cudaFree(didx);
    }
    {
        const float x[1] = {wx}, y[1] = {wy}, p[1] = {90.f}, v[1] = {8.f};
        float *dx = to_dev(x, 1), *dy = to_dev(y, 1), *dp = to_dev(p, 1), *dv = to_dev(v, 1), *dt = to_dev(trk, 3);
        OK(ce2e_tracking_error(h, 0, NULL, dx, dy, dp, dv, 0, dt, 3, 1, NULL));
        to_host(trk, dt, 3);
        CHECK(trk[0] == 0.f && trk[1] == 0.f && trk[2] == 0.f, "tracking error %g %g %g", trk[0], trk[1], trk[2]);
        cudaFree(dx); cudaFree(dy); cudaFree(dp); cudaFree(dv); cudaFree(dt);
    }
    {
        /* one row, one far-away pad vehicle: [ego 6 | tracking 3 | x y v phi] */
        float obs[13] = {5.f, 0.f, 0.f, 0.f, 0.f, 90.f, 0.f, 0.f, -3.f, -45.f, -5.625f, 0.f, 0.f};
        const float act[2] = {0.f, (0.75f / 2.25f)};      /* a_x = 2.25 a - 0.75 ~ 0 */
        float out5[5], nxt[13], sc[2];
        ce2e_turn_classes turn;
        float *dobs, *dact = to_dev(act, 2), *dnext = to_dev(obs, 13), *d5 = to_dev(obs, 5), *dsc = to_dev(act, 2);
        obs[3] = wx; obs[4] = wy;
        dobs = to_dev(obs, 13);
        memset(&turn, 0, sizeof(turn));
        OK(ce2e_rollout_step(h, 0, NULL, dobs, 13, dact, &turn, 1, 1, 0, dnext, 13, d5, dsc, 1, NULL));
        to_host(out5, d5, 5);
        to_host(nxt, dnext, 13);
        to_host(sc, dsc, 2);
        CHECK(out5[3] == 0.f, "veh2veh4real = %g", out5[3]);
        CHECK(out5[4] == 0.f, "veh2road4real = %g (ego in lane)", out5[4]);
        CHECK(out5[1] == 0.f && out5[2] == 0.f, "punish terms %g %g", out5[1], out5[2]);
        /* rewards = 0.05 * -(d_v)^2 + 0.05 * -(a_x)^2 with d_v = -3, a_x ~ 0 (DM:198-207, 297-298) */
        CHECK(fabsf(out5[0] - (-0.45f)) < 1e-6f, "rewards = %.9g", out5[0]);
        CHECK(fabsf(sc[1]) < 1e-7f && sc[0] == 0.f, "scaled action %g %g", sc[0], sc[1]);
        CHECK(fabsf(nxt[4] - (wy + 0.5f)) < 1e-5f && fabsf(nxt[3] - wx) < 1e-6f, "next x y = %g %g", nxt[3], nxt[4]);
        CHECK(nxt[9] == -45.f && nxt[10] == -5.625f && nxt[11] == 0.f && nxt[12] == 0.f,
              "a standing vehicle outside the box does not move: %g %g %g %g", nxt[9], nxt[10], nxt[11], nxt[12]);
        CHECK(fabsf(nxt[6]) < 1e-6f && fabsf(nxt[7]) < 1e-4f && fabsf(nxt[8] - (nxt[0] - 8.f)) < 1e-6f,
              "next tracking %g %g %g", nxt[6], nxt[7], nxt[8]);
        /* error convention */
        CHECK(ce2e_rollout_step(h, 9, NULL, dobs, 13, dact, &turn, 1, 1, 0, dnext, 13, d5, NULL, 1, NULL) == CE2E_ERR_PATH,
              "path_index 9 accepted");
        CHECK(strlen(ce2e_last_error()) > 0, "no error text");
        CHECK(ce2e_rollout_step(h, 0, NULL, NULL, 13, dact, &turn, 1, 1, 0, dnext, 13, d5, NULL, 1, NULL) == CE2E_ERR_NULL,
              "NULL obs accepted");
        CHECK(ce2e_rollout_step(h, 0, NULL, dobs, 13, dact, &turn, 1, 1, 0, dobs, 13, d5, NULL, 1, NULL) == CE2E_ERR_SHAPE,
              "aliased obs accepted");
        CHECK(ce2e_rollout_step(h, 0, NULL, dobs, 13, dact, &turn, 1, 1, 0, dnext, 13, d5, NULL, 0, NULL) == CE2E_OK,
              "empty batch rejected");
        cudaFree(dobs); cudaFree(dact); cudaFree(dnext); cudaFree(d5); cudaFree(dsc);
    }
    OK(ce2e_paths_destroy(h));
}

static void test_errors_without_work(void) {
    float dummy;
    CHECK(ce2e_version() == CE2E_VERSION, "version %d", ce2e_version());
    CHECK(ce2e_compute_rewards(7, &dummy, 9, &dummy, 0, 0, &dummy, NULL, 1, NULL) == CE2E_ERR_TASK, "task 7 accepted");
    CHECK(ce2e_action_transform(NULL, &dummy, 1, NULL) == CE2E_ERR_NULL, "NULL actions accepted");
    CHECK(ce2e_dynamics_step(&dummy, 5, &dummy, 0.1, &dummy, 6, NULL, 0, 1, NULL) == CE2E_ERR_SHAPE, "ld 5 accepted");
}

int main(void) {
    const int64_t n0 = ce2e_launch_count();
    test_errors_without_work();
    test_action_transform();
    test_dynamics();
    test_paths_and_rollout();
    printf("launches %lld\n", (long long)(ce2e_launch_count() - n0));
    printf(failures ? "FAILED %d\n" : "ALL OK%.0d\n", failures);
    return failures ? 1 : 0;
}
