// ce2e.cu -- kernels and C ABI of libce2e.so (see include/ce2e.h).
//
// Build (done by __graft_entry__.build()):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared
//        -Xcompiler -fPIC,-ffp-contract=off -I include -o libce2e.so ce2e.cu
//
// Kernel map (reference citations in ce2e_device.cuh / include/ce2e.h):
//   k_model_step<G>   the fused EnvironmentModel.rollout_out step (also serves
//                     compute_rewards and compute_next_obses through flags)
//   k_dynamics_step   VehicleDynamics.f_xu / prediction / ego_predict
//   k_tracking        ReferencePath.tracking_error_vector
//   k_closest         ReferencePath.find_closest_point (any ratio)
//   k_index_points    indexs2points / future_n_data
//   k_veh_predict     EnvironmentModel.veh_predict
//   k_action          _action_transformation_for_end2end
//   k_ss              EnvironmentModel.ss
#include "ce2e.h"
#include "ce2e_device.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace ce2e;

// ------------------------------------------------------------------------------------------
// host-side state
// ------------------------------------------------------------------------------------------
namespace {

thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CE2E_CUDA(expr)                                                                         \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(CE2E_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                              \
    } while (0)

int after_launch(const char *what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(CE2E_ERR_CUDA, "%s launch: %s", what, cudaGetErrorString(e));
    return CE2E_OK;
}

struct DeviceInfo {
    int sms = 0;
    int max_smem_optin = 0;
    bool ok = false;
};
int device_info(DeviceInfo **out) {
    static thread_local DeviceInfo info[64];
    int dev = 0;
    CE2E_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(CE2E_ERR_CUDA, "device ordinal %d out of range", dev);
    if (!info[dev].ok) {
        CE2E_CUDA(cudaDeviceGetAttribute(&info[dev].sms, cudaDevAttrMultiProcessorCount, dev));
        CE2E_CUDA(cudaDeviceGetAttribute(&info[dev].max_smem_optin,
                                         cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        info[dev].ok = true;
    }
    *out = &info[dev];
    return CE2E_OK;
}

// VehicleDynamics.__init__ constants (DM:38-46) folded in fp32 in source order (appendix A).
DynConsts make_dyn_consts(double tau_py) {
    const float C_f = -155495.0f, C_r = -155495.0f, a = 1.19f, b = 1.46f, mass = 1520.0f,
                I_z = 2642.0f, miu = 0.8f, g = 9.81f;
    const float tau = (float)tau_py;
    DynConsts k;
    k.tau = tau;
    k.m = mass;
    k.Iz = I_z;
    k.a = a;
    k.b = b;
    float aCf = a * C_f, bCr = b * C_r;
    float K1 = aCf - bCr;
    k.tauK1 = tau * K1;
    k.tauCf = tau * C_f;
    k.taum = tau * mass;
    float CfCr = C_f + C_r;
    k.Dv = tau * CfCr;
    float taua = tau * a;
    k.tauaCf = taua * C_f;
    float a2 = a * a, b2 = b * b;
    float a2Cf = a2 * C_f, b2Cr = b2 * C_r;
    float s2 = a2Cf + b2Cr;
    k.Dr = tau * s2;
    float ab = a + b;
    float bm = b * mass, am = a * mass;
    float bmg = bm * g, amg = am * g;
    k.Fzf = bmg / ab;
    k.Fzr = amg / ab;
    k.muFzf = miu * k.Fzf;
    k.muFzr = miu * k.Fzr;
    return k;
}

}  // namespace

struct ce2e_paths {
    int task;
    int n_paths;
    int L[CE2E_MAX_PATHS];
    int N10[CE2E_MAX_PATHS];
    int stride10;                 // even, entries per path in the decimated tables
    float tail[CE2E_MAX_PATHS][3];
    float *full[CE2E_MAX_PATHS];  // device [3, L]: x | y | phi
    float2 *xy10;                 // device [n_paths, stride10]
    float *phi10;                 // device [n_paths, stride10]
    int device;
};

namespace {

PathView make_view(const ce2e_paths *p) {
    PathView v;
    v.xy = p->xy10;
    v.phi = p->phi10;
    v.stride = p->stride10;
    v.n_paths = p->n_paths;
    for (int i = 0; i < 4; ++i) {
        v.N[i] = i < p->n_paths ? p->N10[i] : 0;
        v.L[i] = i < p->n_paths ? p->L[i] : 0;
        for (int j = 0; j < 3; ++j) v.tail[i][j] = i < p->n_paths ? p->tail[i][j] : 0.f;
    }
    return v;
}

// ------------------------------------------------------------------------------------------
// fused model step
// ------------------------------------------------------------------------------------------
constexpr int F_REWARD = 1;      // compute_rewards on obs_in
constexpr int F_NEXT = 2;        // compute_next_obses -> obs_out
constexpr int F_ACT_NORM = 4;    // actions are normalised: apply the action transformation
constexpr int F_VEC_IN = 8;      // vehicle block of obs_in is 16 B aligned (float4 loads)
constexpr int F_VEC_OUT = 16;    // same for obs_out

struct StepParams {
    PathView pv;
    DynConsts dyn;
    const float *obs_in;
    float *obs_out;
    const float *act;
    const int32_t *ref_idx;
    float *out5;
    float *dict16;
    float *act_scaled_out;
    int64_t ld_in, ld_out, B;
    int task, path_index, V_in, V_out, n_future, flags;
    int S;                       // lanes cooperating on one row in the ego phase (1, 2, 4, 8)
    ce2e_turn_classes turn;
};

constexpr int STEP_THREADS = 256;
constexpr int STEP_WARPS = STEP_THREADS / 32;

__device__ __forceinline__ float4 load_veh(const float *p, bool vec) {
    if (vec) return *reinterpret_cast<const float4 *>(p);
    return make_float4(p[0], p[1], p[2], p[3]);
}
__device__ __forceinline__ void store_veh(float *p, float4 v, bool vec) {
    if (vec) {
        *reinterpret_cast<float4 *>(p) = v;
    } else {
        p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
    }
}

// Work decomposition (DESIGN.md "k_model_step"):
//   A warp owns a tile of E = 32/S consecutive rows.
//   Ego phase   : lane -> (row = lane % E, part = lane / E).  Each lane runs the row's scalar
//                 chain (action scaling, reward terms, road terms, f_xu); the S lanes of a row
//                 split the waypoint scan and merge with a first-minimum rule.
//   Vehicle phase: lane -> (row group = lane / G, vehicle = lane % G); the G lanes of a group
//                 read the row's vehicle block with one coalesced (float4) access per lane,
//                 add their hinge terms with an xor-shuffle tree and write the predicted
//                 vehicles back coalesced.
template <int G>
__global__ void __launch_bounds__(STEP_THREADS)
k_model_step(const __grid_constant__ StepParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared layout: float4 ego[WARPS][32] | float2 xy[n_paths*stride] | float phi[n_paths*stride]
    float4 *s_ego = reinterpret_cast<float4 *>(smem_raw);
    float2 *s_xy = reinterpret_cast<float2 *>(s_ego + STEP_WARPS * 32);
    float *s_phi = reinterpret_cast<float *>(s_xy + (size_t)P.pv.n_paths * P.pv.stride);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool do_next = P.flags & F_NEXT, do_rew = P.flags & F_REWARD;
    if (do_next) {
        const int tot = P.pv.n_paths * P.pv.stride;
        for (int i = tid; i < tot; i += STEP_THREADS) {
            s_xy[i] = P.pv.xy[i];
            s_phi[i] = P.pv.phi[i];
        }
        __syncthreads();
    }
    const int S = P.S, E = 32 / S;
    const int el = lane % E, part = lane / E;
    const int n_trk = 3 * (P.n_future + 1);
    const int veh_off = 6 + n_trk;
    constexpr int EPW = 32 / G;              // rows per vehicle-phase pass
    const int sub = lane / G, vl = lane % G;
    const int64_t n_tiles = (P.B + E - 1) / E;
    float4 *my_ego = s_ego + warp * 32;

    for (int64_t tile = (int64_t)blockIdx.x * STEP_WARPS + warp; tile < n_tiles;
         tile += (int64_t)gridDim.x * STEP_WARPS) {
        const int64_t row0 = tile * E;
        const int64_t row = row0 + el;
        const bool valid = row < P.B;
        const int64_t rr = valid ? row : P.B - 1;
        const float *o = P.obs_in + rr * P.ld_in;

        // ---------------- ego phase ----------------
        const float vx = o[0], vy = o[1], r = o[2], x = o[3], y = o[4], phi_deg = o[5];
        float steer = P.act[2 * rr], a_x = P.act[2 * rr + 1];
        if (P.flags & F_ACT_NORM) action_transform(steer, a_x, steer, a_x);
        const float phi = deg2rad(phi_deg);
        float s, c;
        sincosf(phi, &s, &c);

        float rewards = 0.f, v2r_tr = 0.f, v2r_re = 0.f;
        float punish_steer = 0.f, punish_a_x = 0.f, punish_yaw = 0.f, devi_v = 0.f, devi_y = 0.f,
              devi_phi = 0.f;
        if (do_rew) {
            punish_steer = -sq(steer);                                   // DM:198-207
            punish_a_x = -sq(a_x);
            punish_yaw = -sq(r);
            devi_y = -sq(o[6]);
            devi_phi = -sq(deg2rad(o[7]));
            devi_v = -sq(o[8]);
            rewards = ((((0.05f * devi_v + 0.8f * devi_y) + 30.0f * devi_phi) + 0.02f * punish_yaw) +
                       5.0f * punish_steer) + 0.05f * punish_a_x;       // DM:297-298
            Circles ec = circle_centres(x, y, s, c);
            road_terms(P.task, ec.fx, ec.fy, v2r_tr, v2r_re);
            road_terms(P.task, ec.rx, ec.ry, v2r_tr, v2r_re);
            if (part == 0) my_ego[el] = make_float4(ec.fx, ec.fy, ec.rx, ec.ry);
        }

        if (do_next) {
            float nxt[6];
            f_xu_next(P.dyn, vx, vy, r, x, y, phi, s, c, steer, a_x, nxt);
            nxt[0] = fminf(fmaxf(nxt[0], 0.0f), 35.0f);                  // ego_predict, DM:390
            int p = P.ref_idx ? P.ref_idx[rr] : P.path_index;
            const bool p_ok = (p >= 0) && (p < P.pv.n_paths);
            p = p_ok ? p : 0;
            const float2 *t_xy = s_xy + (size_t)p * P.pv.stride;
            // find_closest_point: the S lanes of a row scan disjoint even-aligned chunks
            const int n_even = (P.pv.N[p] + 1) & ~1;
            int chunk = ((n_even + S - 1) / S + 1) & ~1;
            int k0 = min(part * chunk, n_even), k1 = min(k0 + chunk, n_even);
            float best;
            int bi;
            scan_min(t_xy, k0, k1, nxt[3], nxt[4], best, bi);
            for (int off = E; off < 32; off <<= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, best, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (part == 0 && valid) {
                float *q = P.obs_out + row * P.ld_out;
#pragma unroll
                for (int i = 0; i < 6; ++i) q[i] = nxt[i];
                if (p_ok) {
                    tracking_from_index(t_xy, s_phi + (size_t)p * P.pv.stride, P.pv.L[p],
                                        P.pv.tail[p], P.task, bi, nxt[3], nxt[4], nxt[5], nxt[0],
                                        P.n_future, q + 6);
                } else {
                    for (int i = 0; i < n_trk; ++i) q[6 + i] = 0.0f;     // DM:342-343
                }
            }
        }
        if (P.act_scaled_out && part == 0 && valid) {
            P.act_scaled_out[2 * row] = steer;
            P.act_scaled_out[2 * row + 1] = a_x;
        }
        __syncwarp();

        // ---------------- vehicle phase ----------------
        float v2v_tr = 0.f, v2v_re = 0.f;
        if (P.V_in > 0) {
            const bool vec_in = P.flags & F_VEC_IN, vec_out = P.flags & F_VEC_OUT;
            for (int p0 = 0; p0 < E; p0 += EPW) {
                const int e2 = p0 + sub;
                const bool ev = (e2 < E) && (row0 + e2 < P.B);
                float acc_tr = 0.f, acc_re = 0.f;
                if (ev) {
                    const float4 ec = do_rew ? my_ego[e2] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float *vin = P.obs_in + (row0 + e2) * P.ld_in + veh_off;
                    float *vout = do_next ? P.obs_out + (row0 + e2) * P.ld_out + veh_off : nullptr;
                    for (int j = vl; j < P.V_in; j += G) {
                        const float4 v = load_veh(vin + 4 * j, vec_in);
                        const float th = deg2rad(v.w);
                        float vs, vc;
                        sincosf(th, &vs, &vc);
                        if (do_rew) {
                            const Circles vcirc = circle_centres(v.x, v.y, vs, vc);
                            pair_term(ec.x, ec.y, vcirc.fx, vcirc.fy, acc_tr, acc_re);
                            pair_term(ec.x, ec.y, vcirc.rx, vcirc.ry, acc_tr, acc_re);
                            pair_term(ec.z, ec.w, vcirc.fx, vcirc.fy, acc_tr, acc_re);
                            pair_term(ec.z, ec.w, vcirc.rx, vcirc.ry, acc_tr, acc_re);
                        }
                        if (do_next && j < P.V_out)
                            store_veh(vout + 4 * j, veh_predict_one(v, th, vs, vc, P.turn.tc[j]),
                                      vec_out);
                    }
                }
                if (do_rew) {
#pragma unroll
                    for (int off = G / 2; off > 0; off >>= 1) {
                        acc_tr += __shfl_xor_sync(0xffffffffu, acc_tr, off);
                        acc_re += __shfl_xor_sync(0xffffffffu, acc_re, off);
                    }
                    // hand the sums to the lane that owns the row in the ego phase
                    const int src = ((lane - p0) * G) & 31;
                    const float g_tr = __shfl_sync(0xffffffffu, acc_tr, src);
                    const float g_re = __shfl_sync(0xffffffffu, acc_re, src);
                    if (lane >= p0 && lane < p0 + EPW) { v2v_tr = g_tr; v2v_re = g_re; }
                }
            }
        }
        __syncwarp();

        if (do_rew && part == 0 && valid) {
            float *o5 = P.out5;
            o5[row] = rewards;
            o5[P.B + row] = v2v_tr + v2r_tr;                              // DM:299
            o5[2 * P.B + row] = v2v_re + v2r_re;                          // DM:300
            o5[3 * P.B + row] = v2v_re;
            o5[4 * P.B + row] = v2r_re;
            if (P.dict16) {
                float *d = P.dict16 + row;
                const int64_t B = P.B;
                d[0] = punish_steer; d[B] = punish_a_x; d[2 * B] = punish_yaw; d[3 * B] = devi_v;
                d[4 * B] = devi_y; d[5 * B] = devi_phi; d[6 * B] = 5.0f * punish_steer;
                d[7 * B] = 0.05f * punish_a_x; d[8 * B] = 0.02f * punish_yaw;
                d[9 * B] = 0.05f * devi_v; d[10 * B] = 0.8f * devi_y; d[11 * B] = 30.0f * devi_phi;
                d[12 * B] = v2v_tr; d[13 * B] = v2r_tr; d[14 * B] = v2v_re; d[15 * B] = v2r_re;
            }
        }
    }
}

int pick_group(int V) {
    int g = 1;
    while (g < V && g < 32) g <<= 1;
    return g;
}

int launch_model_step(StepParams &P, cudaStream_t st) {
    DeviceInfo *di;
    int rc = device_info(&di);
    if (rc) return rc;
    // lanes per row in the ego phase: keep >= ~16 warps per SM busy when the batch is small
    int S = 1;
    while (S < 8 && (P.B * S + 31) / 32 < (int64_t)di->sms * 16) S <<= 1;
    if (!(P.flags & F_NEXT)) S = 1;
    P.S = S;
    const int E = 32 / S;
    const int64_t n_tiles = (P.B + E - 1) / E;
    size_t smem = (size_t)P.pv.n_paths * P.pv.stride * 12 + STEP_WARPS * 32 * sizeof(float4);
    if ((int)smem > di->max_smem_optin)
        return fail(CE2E_ERR_SHAPE, "path tables need %zu B of shared memory (max %d)", smem,
                    di->max_smem_optin);
    const int G = pick_group(P.V_in);
    void (*kern)(const StepParams) = nullptr;
    switch (G) {
        case 1: kern = k_model_step<1>; break;
        case 2: kern = k_model_step<2>; break;
        case 4: kern = k_model_step<4>; break;
        case 8: kern = k_model_step<8>; break;
        case 16: kern = k_model_step<16>; break;
        default: kern = k_model_step<32>; break;
    }
    if (smem > 48 * 1024)
        CE2E_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (n_tiles + STEP_WARPS - 1) / STEP_WARPS;
    const int64_t max_blocks = (int64_t)di->sms * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    kern<<<(unsigned)blocks, STEP_THREADS, smem, st>>>(P);
    return after_launch("k_model_step");
}

// ------------------------------------------------------------------------------------------
// standalone kernels
// ------------------------------------------------------------------------------------------
__global__ void k_action(const float *__restrict__ in, float *__restrict__ out, int64_t B) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float s, a;
    action_transform(in[2 * i], in[2 * i + 1], s, a);
    out[2 * i] = s;
    out[2 * i + 1] = a;
}

__global__ void k_dynamics_step(const __grid_constant__ DynConsts K, const float *__restrict__ st,
                                int64_t ld_s, const float *__restrict__ act, float *__restrict__ nx,
                                int64_t ld_n, float *__restrict__ params, int clip, int64_t B) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float *o = st + i * ld_s;
    const float vx = o[0], vy = o[1], r = o[2], x = o[3], y = o[4], phi_deg = o[5];
    const float steer = act[2 * i], a_x = act[2 * i + 1];
    const float phi = deg2rad(phi_deg);
    float s, c;
    sincosf(phi, &s, &c);
    float out[6];
    f_xu_next(K, vx, vy, r, x, y, phi, s, c, steer, a_x, out);
    if (clip) out[0] = fminf(fmaxf(out[0], 0.0f), 35.0f);
    float *q = nx + i * ld_n;
#pragma unroll
    for (int k = 0; k < 6; ++k) q[k] = out[k];
    if (params) {
        float pr[4];
        f_xu_params(K, vx, vy, r, steer, a_x, pr);
        *reinterpret_cast<float4 *>(params + 4 * i) = make_float4(pr[0], pr[1], pr[2], pr[3]);
    }
}

// tracking_error_vector: one thread per row, decimated tables read through the read-only path.
__global__ void k_tracking(const __grid_constant__ PathView pv, int task, int path_index,
                           const int32_t *__restrict__ ref_idx, const float *__restrict__ xs,
                           const float *__restrict__ ys, const float *__restrict__ phis,
                           const float *__restrict__ vs, int n_future, float *__restrict__ out,
                           int64_t ld_out, int64_t B) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    int p = ref_idx ? ref_idx[i] : path_index;
    float *q = out + i * ld_out;
    if (p < 0 || p >= pv.n_paths) {
        for (int k = 0; k < 3 * (n_future + 1); ++k) q[k] = 0.0f;
        return;
    }
    const float x = xs[i], y = ys[i];
    float best;
    int bi;
    const float2 *t_xy = pv.xy + (size_t)p * pv.stride;
    scan_min(t_xy, 0, (pv.N[p] + 1) & ~1, x, y, best, bi);
    tracking_from_index(t_xy, pv.phi + (size_t)p * pv.stride, pv.L[p], pv.tail[p], task, bi, x, y,
                        phis[i], vs[i], n_future, q);
}

// find_closest_point with an arbitrary decimation ratio on the full table [3, L].
__global__ void k_closest(const float *__restrict__ full, int L, int ratio,
                          const float *__restrict__ xs, const float *__restrict__ ys,
                          int64_t *__restrict__ idx_out, float *__restrict__ pts_out, int64_t B) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float x = xs[i], y = ys[i];
    const float *px = full, *py = full + L;
    float best = CUDART_INF_F;
    int bi = 0;
    for (int k = 0; k < L; k += ratio) {
        float d = sq(x - px[k]) + sq(y - py[k]);
        if (d < best) { best = d; bi = k; }
    }
    if (idx_out) idx_out[i] = bi;
    if (pts_out) {
        pts_out[i] = px[bi];
        pts_out[B + i] = py[bi];
        pts_out[2 * B + i] = full[2 * (size_t)L + bi];
    }
}

__global__ void k_index_points(const float *__restrict__ full, int L, const int64_t *__restrict__ idx,
                               int n_future, float *__restrict__ pts_out, int64_t B) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (n_future == 0) {
        int64_t k = idx[i];
        k = k >= 0 ? k : 0;
        k = k < L ? k : L - 1;
        pts_out[i] = full[k];
        pts_out[B + i] = full[L + k];
        pts_out[2 * B + i] = full[2 * (size_t)L + k];
        return;
    }
    int k = (int)idx[i];                                   // tf.cast(current_indexs, tf.int32), DM:719
    for (int f = 0; f < n_future; ++f) {
        k += 80;
        if (k >= L - 2) k = L - 2;
        int kk = k >= 0 ? k : 0;
        kk = kk < L ? kk : L - 1;
        float *q = pts_out + (size_t)f * 3 * B;
        q[i] = full[kk];
        q[B + i] = full[L + kk];
        q[2 * B + i] = full[2 * (size_t)L + kk];
    }
}

__global__ void k_veh_predict(const float *__restrict__ vin, int64_t ld_in,
                              const __grid_constant__ ce2e_turn_classes turn, int V,
                              float *__restrict__ vout, int64_t ld_out, int64_t B) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * V) return;
    const int64_t i = t / V;
    const int j = (int)(t - i * V);
    const float *p = vin + i * ld_in + 4 * j;
    const float4 v = make_float4(p[0], p[1], p[2], p[3]);
    const float th = deg2rad(v.w);
    float s, c;
    sincosf(th, &s, &c);
    const float4 n = veh_predict_one(v, th, s, c, turn.tc[j]);
    float *q = vout + i * ld_out + 4 * j;
    q[0] = n.x; q[1] = n.y; q[2] = n.z; q[3] = n.w;
}

// EnvironmentModel.ss (DM:134-184): one thread per row, vehicles in source order.
__global__ void k_ss(const float *__restrict__ obs, int64_t ld, const float *__restrict__ nobs,
                     int64_t ldn, int V, int veh_off, float one_m_lam, float *__restrict__ out,
                     int64_t B) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float *o = obs + i * ld, *n = nobs + i * ldn;
    float s, c;
    sincosf(deg2rad(o[5]), &s, &c);
    const Circles e0 = circle_centres(o[3], o[4], s, c);
    sincosf(deg2rad(n[5]), &s, &c);
    const Circles e1 = circle_centres(n[3], n[4], s, c);
    float acc = 0.f;
    for (int j = 0; j < V; ++j) {
        const float *v = o + veh_off + 4 * j, *w = n + veh_off + 4 * j;
        const float ego2veh = __fsqrt_rn(sq(o[3] - v[0]) + sq(o[4] - v[1]));
        sincosf(deg2rad(v[3]), &s, &c);
        const Circles v0 = circle_centres(v[0], v[1], s, c);
        sincosf(deg2rad(w[3]), &s, &c);
        const Circles v1 = circle_centres(w[0], w[1], s, c);
        const float ex0[2] = {e0.fx, e0.rx}, ey0[2] = {e0.fy, e0.ry}, ex1[2] = {e1.fx, e1.rx},
                    ey1[2] = {e1.fy, e1.ry};
        const float vx0[2] = {v0.fx, v0.rx}, vy0[2] = {v0.fy, v0.ry}, vx1[2] = {v1.fx, v1.rx},
                    vy1[2] = {v1.fy, v1.ry};
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const float d = __fsqrt_rn(sq(ex0[a] - vx0[b]) + sq(ey0[a] - vy0[b]));
                const float nd = __fsqrt_rn(sq(ex1[a] - vx1[b]) + sq(ey1[a] - vy1[b]));
                const float h = (nd - 2.5f) - one_m_lam * (d - 2.5f);
                acc = acc + ((h < 0.0f && ego2veh < 10.0f) ? sq(h) : 0.0f);
            }
    }
    out[i] = acc;
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

int check_task(int task) {
    if (task < 0 || task > 2) return fail(CE2E_ERR_TASK, "task %d not in {0:left, 1:straight, 2:right}", task);
    return CE2E_OK;
}

int check_batch(int64_t B) {
    if (B < 0 || B > ((int64_t)1 << 40)) return fail(CE2E_ERR_SHAPE, "bad batch size %lld", (long long)B);
    return CE2E_OK;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int model_step_common(const ce2e_paths *paths, int task, int path_index, const int32_t *ref_idx,
                      const float *obs_in, int64_t ld_in, const float *act,
                      const ce2e_turn_classes *turn, int V_in, int V_out, int n_future,
                      float *obs_out, int64_t ld_out, float *out5, float *dict16,
                      float *act_scaled_out, int64_t B, int flags, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if ((rc = check_task(task))) return rc;
    if (V_in < 0 || V_in > CE2E_MAX_VEH || V_out < 0 || V_out > V_in)
        return fail(CE2E_ERR_SHAPE, "bad vehicle counts V_in=%d V_out=%d (max %d)", V_in, V_out,
                    CE2E_MAX_VEH);
    if (n_future < 0 || n_future > 1024) return fail(CE2E_ERR_SHAPE, "bad num_future_data %d", n_future);
    const int D_in = 6 + 3 * (n_future + 1) + 4 * V_in, D_out = 6 + 3 * (n_future + 1) + 4 * V_out;
    if (B == 0) return CE2E_OK;                   // empty batch: nothing to read or write
    if (!obs_in || !act) return fail(CE2E_ERR_NULL, "obs_in / actions is NULL");
    if (ld_in < D_in) return fail(CE2E_ERR_SHAPE, "ld_in=%lld < D=%d", (long long)ld_in, D_in);
    StepParams P;
    memset(&P, 0, sizeof(P));
    if (flags & F_REWARD) {
        if (!out5) return fail(CE2E_ERR_NULL, "out5 is NULL");
    }
    if (flags & F_NEXT) {
        if (!paths) return fail(CE2E_ERR_NULL, "paths handle is NULL");
        if (!obs_out) return fail(CE2E_ERR_NULL, "obs_out is NULL");
        if (ld_out < D_out) return fail(CE2E_ERR_SHAPE, "ld_out=%lld < D=%d", (long long)ld_out, D_out);
        if (V_out > 0 && !turn) return fail(CE2E_ERR_NULL, "turn classes are NULL");
        if (!ref_idx && (path_index < 0 || path_index >= paths->n_paths))
            return fail(CE2E_ERR_PATH, "path_index %d outside [0, %d)", path_index, paths->n_paths);
        if (obs_out == obs_in) return fail(CE2E_ERR_SHAPE, "obs_out must not alias obs_in");
        P.pv = make_view(paths);
        if (turn) P.turn = *turn;
    } else {
        P.pv.n_paths = 0;
        P.pv.stride = 0;
    }
    P.dyn = make_dyn_consts(1.0 / 10.0);          // prediction(..., base_frequency = 10.), DM:387
    P.obs_in = obs_in; P.obs_out = obs_out; P.act = act; P.ref_idx = ref_idx; P.out5 = out5;
    P.dict16 = dict16; P.act_scaled_out = act_scaled_out;
    P.ld_in = ld_in; P.ld_out = ld_out; P.B = B;
    P.task = task; P.path_index = path_index; P.V_in = V_in; P.V_out = V_out; P.n_future = n_future;
    const int veh_off = 6 + 3 * (n_future + 1);
    if (aligned16(obs_in + veh_off) && ld_in % 4 == 0) flags |= F_VEC_IN;
    if (obs_out && aligned16(obs_out + veh_off) && ld_out % 4 == 0) flags |= F_VEC_OUT;
    P.flags = flags;
    return launch_model_step(P, (cudaStream_t)stream);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int ce2e_version(void) { return CE2E_VERSION; }
const char *ce2e_last_error(void) { return g_err; }
int64_t ce2e_launch_count(void) { return g_launches; }

int ce2e_paths_create(int task, int n_paths, const int32_t *lens, const float *const *xs,
                      const float *const *ys, const float *const *phis, ce2e_paths **out) {
    int rc;
    if ((rc = check_task(task))) return rc;
    if (!lens || !xs || !ys || !phis || !out) return fail(CE2E_ERR_NULL, "NULL argument");
    if (n_paths < 1 || n_paths > CE2E_MAX_PATHS)
        return fail(CE2E_ERR_SHAPE, "n_paths=%d outside [1, %d]", n_paths, CE2E_MAX_PATHS);
    int maxN = 0;
    for (int i = 0; i < n_paths; ++i) {
        if (lens[i] < 3 || lens[i] > (1 << 24)) return fail(CE2E_ERR_SHAPE, "path %d length %d", i, lens[i]);
        if (!xs[i] || !ys[i] || !phis[i]) return fail(CE2E_ERR_NULL, "path %d table is NULL", i);
        int n10 = (lens[i] + 9) / 10;
        if (n10 > maxN) maxN = n10;
    }
    ce2e_paths *h = (ce2e_paths *)calloc(1, sizeof(ce2e_paths));
    if (!h) return fail(CE2E_ERR_NOMEM, "calloc failed");
    h->task = task;
    h->n_paths = n_paths;
    // even stride, and stride*8 B = 16 (mod 128) so that rows of different paths read by the
    // lanes of one warp (per-row ref_idx) fall into different shared-memory banks
    int stride = (maxN + 1) & ~1;
    while (stride % 16 != 2) stride += 2;
    h->stride10 = stride;
    cudaGetDevice(&h->device);
    std::vector<float2> xy((size_t)n_paths * stride, make_float2(1e30f, 1e30f));
    std::vector<float> ph((size_t)n_paths * stride, 0.f);
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n_paths && e == cudaSuccess; ++i) {
        const int L = lens[i];
        h->L[i] = L;
        h->N10[i] = (L + 9) / 10;
        for (int k = 0; k < h->N10[i]; ++k) {
            xy[(size_t)i * stride + k] = make_float2(xs[i][10 * k], ys[i][10 * k]);
            ph[(size_t)i * stride + k] = phis[i][10 * k];
        }
        h->tail[i][0] = xs[i][L - 2];
        h->tail[i][1] = ys[i][L - 2];
        h->tail[i][2] = phis[i][L - 2];
        e = cudaMalloc((void **)&h->full[i], sizeof(float) * 3 * (size_t)L);
        if (e == cudaSuccess) e = cudaMemcpy(h->full[i], xs[i], sizeof(float) * L, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->full[i] + L, ys[i], sizeof(float) * L, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(h->full[i] + 2 * (size_t)L, phis[i], sizeof(float) * L, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->xy10, sizeof(float2) * xy.size());
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->phi10, sizeof(float) * ph.size());
    if (e == cudaSuccess) e = cudaMemcpy(h->xy10, xy.data(), sizeof(float2) * xy.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->phi10, ph.data(), sizeof(float) * ph.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        ce2e_paths_destroy(h);
        return fail(CE2E_ERR_CUDA, "ce2e_paths_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return CE2E_OK;
}

int ce2e_paths_destroy(ce2e_paths *h) {
    if (!h) return CE2E_OK;
    for (int i = 0; i < CE2E_MAX_PATHS; ++i)
        if (h->full[i]) cudaFree(h->full[i]);
    if (h->xy10) cudaFree(h->xy10);
    if (h->phi10) cudaFree(h->phi10);
    free(h);
    return CE2E_OK;
}

int ce2e_action_transform(const float *act_norm, float *act_scaled, int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!act_norm || !act_scaled) return fail(CE2E_ERR_NULL, "NULL argument");
    k_action<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(act_norm, act_scaled, B);
    return after_launch("k_action");
}

int ce2e_dynamics_step(const float *states, int64_t ld_states, const float *actions, double tau,
                       float *next, int64_t ld_next, float *params, int clip_vx, int64_t B,
                       void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!states || !actions || !next) return fail(CE2E_ERR_NULL, "NULL argument");
    if (ld_states < 6 || ld_next < 6) return fail(CE2E_ERR_SHAPE, "ld < 6");
    if (params && !aligned16(params)) return fail(CE2E_ERR_SHAPE, "params must be 16 B aligned");
    const DynConsts K = make_dyn_consts(tau);
    k_dynamics_step<<<blocks_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
        K, states, ld_states, actions, next, ld_next, params, clip_vx, B);
    return after_launch("k_dynamics_step");
}

int ce2e_find_closest_point(const ce2e_paths *paths, int path_index, const float *xs,
                            const float *ys, int ratio, int64_t *idx_out, float *pts_out,
                            int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!paths || !xs || !ys) return fail(CE2E_ERR_NULL, "NULL argument");
    if (path_index < 0 || path_index >= paths->n_paths)
        return fail(CE2E_ERR_PATH, "path_index %d outside [0, %d)", path_index, paths->n_paths);
    if (ratio < 1) return fail(CE2E_ERR_SHAPE, "ratio %d < 1", ratio);
    k_closest<<<blocks_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
        paths->full[path_index], paths->L[path_index], ratio, xs, ys, idx_out, pts_out, B);
    return after_launch("k_closest");
}

int ce2e_index_points(const ce2e_paths *paths, int path_index, const int64_t *idx, int n_future,
                      float *pts_out, int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!paths || !idx || !pts_out) return fail(CE2E_ERR_NULL, "NULL argument");
    if (path_index < 0 || path_index >= paths->n_paths)
        return fail(CE2E_ERR_PATH, "path_index %d outside [0, %d)", path_index, paths->n_paths);
    if (n_future < 0) return fail(CE2E_ERR_SHAPE, "n_future < 0");
    k_index_points<<<blocks_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
        paths->full[path_index], paths->L[path_index], idx, n_future, pts_out, B);
    return after_launch("k_index_points");
}

int ce2e_tracking_error(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                        const float *xs, const float *ys, const float *phis, const float *vs,
                        int n_future, float *out, int64_t ld_out, int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!paths || !xs || !ys || !phis || !vs || !out) return fail(CE2E_ERR_NULL, "NULL argument");
    if (!ref_idx && (path_index < 0 || path_index >= paths->n_paths))
        return fail(CE2E_ERR_PATH, "path_index %d outside [0, %d)", path_index, paths->n_paths);
    if (n_future < 0 || ld_out < 3 * (n_future + 1)) return fail(CE2E_ERR_SHAPE, "bad n_future / ld_out");
    k_tracking<<<blocks_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
        make_view(paths), paths->task, path_index, ref_idx, xs, ys, phis, vs, n_future, out, ld_out, B);
    return after_launch("k_tracking");
}

int ce2e_compute_rewards(int task, const float *obs, int64_t ld, const float *actions, int V,
                         int n_future, float *out5, float *dict16, int64_t B, void *stream) {
    return model_step_common(nullptr, task, 0, nullptr, obs, ld, actions, nullptr, V, 0, n_future,
                             nullptr, 0, out5, dict16, nullptr, B, F_REWARD, stream);
}

int ce2e_compute_next_obses(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                            const float *obs_in, int64_t ld_in, const float *actions,
                            const ce2e_turn_classes *turn, int V_in, int V_out, int n_future,
                            float *obs_out, int64_t ld_out, int64_t B, void *stream) {
    if (!paths) return fail(CE2E_ERR_NULL, "paths handle is NULL");
    return model_step_common(paths, paths->task, path_index, ref_idx, obs_in, ld_in, actions, turn,
                             V_in, V_out, n_future, obs_out, ld_out, nullptr, nullptr, nullptr, B,
                             F_NEXT, stream);
}

int ce2e_rollout_step(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                      const float *obs_in, int64_t ld_in, const float *act_norm,
                      const ce2e_turn_classes *turn, int V_in, int V_out, int n_future,
                      float *obs_out, int64_t ld_out, float *out5, float *act_scaled_out,
                      int64_t B, void *stream) {
    if (!paths) return fail(CE2E_ERR_NULL, "paths handle is NULL");
    return model_step_common(paths, paths->task, path_index, ref_idx, obs_in, ld_in, act_norm, turn,
                             V_in, V_out, n_future, obs_out, ld_out, out5, nullptr, act_scaled_out,
                             B, F_REWARD | F_NEXT | F_ACT_NORM, stream);
}

int ce2e_veh_predict(const float *veh_in, int64_t ld_in, const ce2e_turn_classes *turn, int V,
                     float *veh_out, int64_t ld_out, int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!veh_in || !veh_out || !turn) return fail(CE2E_ERR_NULL, "NULL argument");
    if (V < 0 || V > CE2E_MAX_VEH || ld_in < 4 * V || ld_out < 4 * V)
        return fail(CE2E_ERR_SHAPE, "bad V=%d / ld", V);
    if (V == 0) return CE2E_OK;
    k_veh_predict<<<blocks_for(B * V, 256), 256, 0, (cudaStream_t)stream>>>(veh_in, ld_in, *turn, V,
                                                                             veh_out, ld_out, B);
    return after_launch("k_veh_predict");
}

int ce2e_ss(const float *obs, int64_t ld, const float *next_obs, int64_t ld_next, int V,
            int n_future, double lam, float *out, int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!obs || !next_obs || !out) return fail(CE2E_ERR_NULL, "NULL argument");
    const int D = 6 + 3 * (n_future + 1) + 4 * V;
    if (V < 0 || n_future < 0 || ld < D || ld_next < D) return fail(CE2E_ERR_SHAPE, "bad V / n_future / ld");
    k_ss<<<blocks_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
        obs, ld, next_obs, ld_next, V, 6 + 3 * (n_future + 1), (float)(1.0 - lam), out, B);
    return after_launch("k_ss");
}

}  // extern "C"
