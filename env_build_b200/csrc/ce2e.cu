// ce2e.cu -- kernels and C ABI of libce2e.so (see include/ce2e.h).
//
// Build (done by __graft_entry__.build()):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared
//        -Xcompiler -fPIC,-ffp-contract=off -I include -o libce2e.so ce2e.cu
//
// Kernel map (reference citations in ce2e_device.cuh / include/ce2e.h):
//   k_model_step      the fused EnvironmentModel.rollout_out step (also serves
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
#include "ce2e_grid.h"
#include "ce2e_rng.h"

#include <atomic>
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
ce2e_config g_config = {4.8, 2.0, 3.75, 3, 50.0, 8.0, 0.05, 0.8, 30.0, 0.02, 5.0, 0.05};
thread_local int g_last_step_kernel = 0;   // 1: k_model_step (cp.async), 2: k_model_step_pair (TMA)
bool g_use_pdl = getenv("CE2E_NO_PDL") == nullptr;
bool g_fast_trig = false;
// A/B switch: ego columns through TMA boxes too (bit 0: loads, bit 1: stores); CE2E_EGO_TMA=0..3
int g_ego_tma = getenv("CE2E_NO_EGO_TMA") ? 0 : (getenv("CE2E_EGO_TMA") ? atoi(getenv("CE2E_EGO_TMA")) : 3);
// warp-pair TMA kernel for the fused step (ce2e_set_tma): 0 off, 1 on (work split chosen by batch size),
// 2 / 3 on with the overlapped / balanced split forced
int g_use_tma = getenv("CE2E_NO_TMA") ? 0 : 1;

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
    uint32_t *cells;              // device: candidate grids of all paths, concatenated
    int cell_off[CE2E_MAX_PATHS];
    GridSpec grid[CE2E_MAX_PATHS];    // nx == 0: no grid
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
constexpr int F_GYM_EGO = 32;    // CrossroadEnd2end._get_next_ego_state post-ops (E2E:281-282) instead of DM:390

// Candidate grid of find_closest_point as the kernels see it (ce2e_grid.h).
struct GridView {
    const uint32_t *cells;   // all paths' cell tables, concatenated
    int off[4];              // first cell of path p
    int nx[4], ny[4];        // nx == 0: no grid for this path (always scan everything)
    float x0[4], y0[4];
    float inv_h;
};

struct StepParams {
    PathView pv;
    GridView gv;
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
    int horizon;                 // FUSED kernels: steps per launch (act = tape [H,B,2], out5 = [H,5,B])
    // per-vehicle turn tables derived from ce2e_turn_classes (DM:416-421): signed arc radius
    // (+26.875 left-turn modes, -15.625 right-turn modes), its rounded reciprocal, and the half
    // width of the box inside which the heading turns (25, or -1 = never for straight modes)
    float turn_rs[CE2E_MAX_VEH], turn_rr[CE2E_MAX_VEH], turn_half[CE2E_MAX_VEH];
};

constexpr int STEP_WARPS = 14;           // two 448-thread blocks per SM = 28 warps (<= 72 registers)
constexpr int STEP_THREADS = STEP_WARPS * 32;
constexpr int RPW = 16;                  // rows per warp tile: two lanes per row
constexpr int VPL = 4;                   // vehicles per lane and staged chunk (16 B each)
constexpr int QCAP = 24;                 // hinge queue entries per lane; flushed before it can overflow

constexpr int FUSED_MAX_CHUNKS = 4;      // horizon-fused mode keeps all vehicle chunks resident: V <= 32
template <int NBUF>
struct WarpScratchT {
    // vehicle chunk buffers (2: double buffering; FUSED_MAX_CHUNKS: the whole row, resident across steps):
    // lane l owns floats [16 l, 16 l + 16); its vehicle e sits at float4 index e ^ ((l >> 1) & 3)
    // (swizzle: lane-strided LDS.128 without bank conflicts)
    float vbuf[NBUF][32 * 4 * VPL];
    float queue[QCAP * 32];              // [entry][lane]: squared distances inside the 3.5 m gate
};
typedef WarpScratchT<2> WarpScratch;

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// mbarrier + 1-D bulk async copy (the TMA path without a tensor map): one thread arms the barrier
// with the byte count and issues the copies; any thread may then poll the phase.
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(unsigned smem_dst, const void *gsrc, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned mbar, unsigned parity) {
    unsigned done;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(mbar), "r"(parity)
                 : "memory");
    return done != 0;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v));
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
    return v;
}

// Squared centre distance of one circle pair (DM:225); pairs inside the 3.5 m gate are queued
// (shared-memory byte address `qa`, 128 B per entry) so that the sqrt / hinge arithmetic runs
// later, densely, in the reference's accumulation order.
__device__ __forceinline__ void pair_gate(float ex, float ey, float px, float py, unsigned &qa) {
    const float dd = sq(ex - px) + sq(ey - py);
    if (dd < 12.25f) {
        sts_f32(qa, dd);
        qa += 128;
    }
}

// One surrounding vehicle of one row: gate its four circle pairs against the ego circles and
// return its predicted state (DM:218-229, DM:405-427).  Branch free.
// predict_for_a_mode (DM:405-427) with the turn class folded into per-vehicle constants: the
// signed division reproduces -(v/R)/10 exactly ((-q) and q round alike), half = -1 disables the arc.
__device__ __forceinline__ float4 veh_predict_tab(float4 v, float th, float s, float c, float rs, float rr,
                                                  float half) {
    const float step = div10(v.z);
    float4 n;
    n.x = v.x + step * c;
    n.y = v.y + step * s;
    n.z = v.z;
    const bool inside = (fabsf(v.x) < half) && (fabsf(v.y) < half);
    const float q = div10(div_const(v.z, rs, rr));
    float t2 = th + (inside ? q : 0.0f);
    t2 = (t2 > CE2E_PI32) ? t2 - CE2E_TWO_PI32 : t2;
    t2 = (t2 <= -CE2E_PI32) ? t2 + CE2E_TWO_PI32 : t2;
    n.w = rad2deg(t2);
    return n;
}

template <bool REW, bool NEXT, bool FAST>
__device__ __forceinline__ float4 vehicle_step(float4 v, const Circles &ec, float rs, float rr, float half,
                                               unsigned &qa) {
    const float th = deg2rad(v.w);
    float vs, vc;
    if (FAST) sincos_mufu(th, vs, vc);
    else sincos_cw(th, vs, vc);
    if (REW) {
        const Circles w = circle_centres(v.x, v.y, vs, vc);
        pair_gate(ec.fx, ec.fy, w.fx, w.fy, qa);
        pair_gate(ec.fx, ec.fy, w.rx, w.ry, qa);
        pair_gate(ec.rx, ec.ry, w.fx, w.fy, qa);
        pair_gate(ec.rx, ec.ry, w.rx, w.ry, qa);
    }
    return NEXT ? veh_predict_tab(v, th, vs, vc, rs, rr, half) : v;
}

// The nearest-waypoint candidate range of (x, y) on path p: the grid cell's [lo, hi] widened to
// even bounds, or everything when the point is off the grid / not finite.
__device__ __forceinline__ void candidate_range(const GridView &gv, int p, int n_even, float x, float y,
                                                int &k0, int &k1) {
    k0 = 0;
    k1 = n_even;
    const float tx = (x - gv.x0[p]) * gv.inv_h, ty = (y - gv.y0[p]) * gv.inv_h;
    if (tx >= 0.0f && ty >= 0.0f && tx < (float)gv.nx[p] && ty < (float)gv.ny[p]) {
        const uint32_t c = __ldg(gv.cells + gv.off[p] + (int)ty * gv.nx[p] + (int)tx);
        k0 = (int)(c & 0xffffu) & ~1;
        k1 = ((int)(c >> 16) + 2) & ~1;
    }
}

// Work decomposition (DESIGN.md "k_model_step"): two lanes per observation row.
//   A warp owns tiles of RPW = 16 consecutive rows; lane -> (row = lane / 2, half h = lane % 2).
//   Ego phase    : both lanes load the row's ego columns and action and take sin/cos of the
//                  heading; then lane h = 0 evaluates the reward and road terms (DM:198-207,
//                  DM:231-298) while lane h = 1 integrates f_xu, finds the closest waypoint in the
//                  row's candidate range and writes the next ego + tracking columns (DM:322-353).
//   Vehicle phase: the vehicle list is split in two halves, lane h owns half h.  The warp streams
//                  its rows' vehicle blocks through shared memory, four vehicles per lane at a
//                  time, with 16 B cp.async copies (4 lanes move 64 contiguous bytes; double
//                  buffered); a lane updates its vehicles in place, two at a time, and the chunk
//                  goes back with coalesced 16 B stores.  Circle pairs inside the 3.5 m gate are
//                  queued per lane and finished (sqrt, both hinge^2 terms) in a dense loop at the
//                  end of the tile, in the reference's order (vehicle, ego circle, vehicle circle).
//                  veh2veh = (sum over the first half) + (sum over the second half): a fixed
//                  order, independent of the batch size.
// FUSED = true (ce2e_rollout_horizon): the warp keeps its tile's whole observation state on chip
// (vehicles in FUSED_MAX_CHUNKS shared-memory chunk buffers, ego + tracking columns in registers)
// and runs P.horizon steps of an open-loop action tape back to back; only the actions are read and
// the five outputs written per step, the observations once at the start / end.  Same arithmetic
// in the same order as `horizon` separate launches.
template <bool REW, bool NEXT, bool FAST = false, bool FUSED = false>
__global__ void __launch_bounds__(STEP_THREADS, FUSED ? 1 : 2)
k_model_step(const __grid_constant__ StepParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared layout: WarpScratch[WARPS] | float2 xy[n_paths*stride] | float phi[n_paths*stride]
    typedef WarpScratchT<FUSED ? FUSED_MAX_CHUNKS : 2> WarpScratch;
    WarpScratch *s_scr = reinterpret_cast<WarpScratch *>(smem_raw);
    float2 *s_xy = reinterpret_cast<float2 *>(s_scr + STEP_WARPS);
    float *s_phi = reinterpret_cast<float *>(s_xy + (size_t)P.pv.n_paths * P.pv.stride);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool vec_in = P.flags & F_VEC_IN, vec_out = P.flags & F_VEC_OUT;
    bool tables_pending = false;
    unsigned s_mbar = 0;
    if (NEXT) {
        // path tables -> shared memory with two bulk copies issued by one thread and tracked by an
        // mbarrier; they are static, so this may run before the previous launch has finished
        // (programmatic dependent launch); a warp polls the mbarrier right before its first
        // waypoint scan, without meeting the other warps of the block
        const int tot = P.pv.n_paths * P.pv.stride;           // stride is even: tot * 8 B is a multiple of 16
        const int tot4 = (tot + 3) & ~3;                      // the phi table is padded to 16 B
        const unsigned sx = (unsigned)__cvta_generic_to_shared(s_xy), sp = (unsigned)__cvta_generic_to_shared(s_phi);
        s_mbar = sp + 4u * tot4;
        if (tid == 0) mbar_init(s_mbar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(s_mbar, 8u * tot + 4u * tot4);
            bulk_copy_g2s(sx, P.pv.xy, 8u * tot, s_mbar);
            bulk_copy_g2s(sp, P.pv.phi, 4u * tot4, s_mbar);
        }
        tables_pending = true;
    }
    // everything below reads what the previous launch of a rollout wrote
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    WarpScratch &scr = s_scr[warp];
    const int lr = lane >> 1, h = lane & 1;
    const unsigned q_lane = (unsigned)__cvta_generic_to_shared(scr.queue + lane);
    const int swz = lr & 3;                                 // this lane's slot swizzle

    const int n_trk = 3 * (P.n_future + 1);
    const int veh_off = 6 + n_trk;
    const int64_t n_tiles = (P.B + RPW - 1) / RPW;
    // vehicle halves: lane h = 0 owns vehicles [0, H), lane h = 1 owns [H, V)
    const int H = (P.V_in + 1) >> 1;
    const int n_chunks = (H + VPL - 1) / VPL;
    // staging geometry: piece i (0..3) of a lane = row (lane / 8) + 4 i of the tile, vehicle
    // (lane % 4) of half (lane / 4) % 2 of the chunk; it lands in the slot of lane
    // 2*row + half at float4 index (vehicle) ^ (row % 4)
    const int p_row = lane >> 3, p_half = (lane >> 2) & 1, p_e = lane & 3;
    const int p_veh = p_half * H + p_e;                     // + VPL * chunk
    const int p_soff = (2 * p_row + p_half) * (4 * VPL) + ((p_e ^ (p_row & 3)) << 2);
    const unsigned s_stage = (unsigned)__cvta_generic_to_shared(scr.vbuf[0] + p_soff);
    constexpr int PIECE_STRIDE = 8 * 4 * VPL;               // floats between a lane's pieces (4 rows)
    const int ld_in = (int)P.ld_in, ld_out = (int)P.ld_out;

    // tile -> (block, warp): consecutive tiles go to different blocks, so every SM gets the same
    // number of tiles up to one
    for (int64_t tile = (int64_t)warp * gridDim.x + blockIdx.x; tile < n_tiles;
         tile += (int64_t)gridDim.x * STEP_WARPS) {
        const int64_t row0 = tile * RPW;
        const int64_t row = row0 + lr;
        const bool valid = row < P.B;
        const int64_t rr = valid ? row : P.B - 1;
        const float *o = P.obs_in + rr * P.ld_in;
        const int rows_here = (int)min((int64_t)RPW, P.B - row0);
        const float *g_in = P.obs_in + row0 * P.ld_in + (p_row * ld_in + veh_off + 4 * p_veh);
        float *g_out = NEXT ? P.obs_out + row0 * P.ld_out + (p_row * ld_out + veh_off + 4 * p_veh) : nullptr;

        auto stage = [&](int ch, int b) {
            const float *src = g_in + ch * (4 * VPL);
            const unsigned dst = s_stage + (unsigned)(b * 32 * 4 * VPL * 4);
            if (vec_in && rows_here == RPW && H + (ch + 1) * VPL <= P.V_in) {
#pragma unroll
                for (int i = 0; i < 4; ++i) cp_async16(dst + (unsigned)(i * PIECE_STRIDE * 4), src + 4 * i * ld_in);
            } else if (p_veh + ch * VPL < (p_half ? P.V_in : min(H, P.V_in))) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (p_row + 4 * i < rows_here) {
                        const float *s2 = src + 4 * i * ld_in;
                        if (vec_in) {
                            cp_async16(dst + (unsigned)(i * PIECE_STRIDE * 4), s2);
                        } else {
                            float *d2 = scr.vbuf[b] + p_soff + i * PIECE_STRIDE;
                            d2[0] = s2[0]; d2[1] = s2[1]; d2[2] = s2[2]; d2[3] = s2[3];
                        }
                    }
                }
            }
            cp_async_commit();
        };
        // ---------------- ego phase ----------------
        float e9[9];
        if (vec_in && veh_off == 9) {           // o[1] is 16 B aligned: 1 scalar + 2 vector loads
            e9[0] = o[0];
            const float4 a = *reinterpret_cast<const float4 *>(o + 1);
            const float4 b = *reinterpret_cast<const float4 *>(o + 5);
            e9[1] = a.x; e9[2] = a.y; e9[3] = a.z; e9[4] = a.w;
            e9[5] = b.x; e9[6] = b.y; e9[7] = b.z; e9[8] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 9; ++i) e9[i] = o[i];
        }
        const float act0_pre = P.act[2 * rr], act1_pre = P.act[2 * rr + 1];
        if (FUSED) {
            for (int ch = 0; ch < n_chunks; ++ch) stage(ch, ch);      // the whole row, once
        } else if (n_chunks > 0) {
            stage(0, 0);
        }
        const int n_steps = FUSED ? P.horizon : 1;
        for (int step = 0; step < n_steps; ++step) {
        const float *act_t = P.act + (FUSED ? (int64_t)step * 2 * P.B : 0);
        const bool last_step = !FUSED || step == n_steps - 1;
        const float vx = e9[0], vy = e9[1], r = e9[2], x = e9[3], y = e9[4], phi_deg = e9[5];
        float steer = (FUSED && step > 0) ? act_t[2 * rr] : act0_pre, a_x = (FUSED && step > 0) ? act_t[2 * rr + 1] : act1_pre;
        if (P.flags & F_ACT_NORM) action_transform(steer, a_x, steer, a_x);
        const float phi = deg2rad(phi_deg);
        float s, c;
        sincos_cw(phi, s, c);
        const Circles ec = circle_centres(x, y, s, c);
        if (tables_pending) {                 // first tile of this warp: the path tables must have landed
            while (!mbar_try_wait(s_mbar, 0)) {}
            tables_pending = false;
        }

        float rewards = 0.f, v2r_tr = 0.f, v2r_re = 0.f;
        if (REW && (h == 0 || !NEXT)) {                                  // reward lane
            const float punish_steer = -sq(steer);                       // DM:198-207
            const float punish_a_x = -sq(a_x);
            const float punish_yaw = -sq(r);
            const float devi_y = -sq(e9[6]);
            const float devi_phi = -sq(deg2rad(e9[7]));
            const float devi_v = -sq(e9[8]);
            rewards = ((((CE2E_K.w_v * devi_v + CE2E_K.w_y * devi_y) + CE2E_K.w_phi * devi_phi) + CE2E_K.w_yaw * punish_yaw) +
                       CE2E_K.w_steer * punish_steer) + CE2E_K.w_ax * punish_a_x;       // DM:297-298
            road_terms(P.task, ec.fx, ec.fy, v2r_tr, v2r_re);
            road_terms(P.task, ec.rx, ec.ry, v2r_tr, v2r_re);
            if (P.dict16 && valid && h == 0) {
                float *d = P.dict16 + row;
                const int64_t B = P.B;
                d[0] = punish_steer; d[B] = punish_a_x; d[2 * B] = punish_yaw; d[3 * B] = devi_v;
                d[4 * B] = devi_y; d[5 * B] = devi_phi; d[6 * B] = CE2E_K.w_steer * punish_steer;
                d[7 * B] = CE2E_K.w_ax * punish_a_x; d[8 * B] = CE2E_K.w_yaw * punish_yaw;
                d[9 * B] = CE2E_K.w_v * devi_v; d[10 * B] = CE2E_K.w_y * devi_y; d[11 * B] = CE2E_K.w_phi * devi_phi;
            }
        }
        float e9n[9];
        if (NEXT && (h == 1 || !REW)) {                                  // dynamics lane
            float nxt[6];
            f_xu_next(P.dyn, vx, vy, r, x, y, phi, s, c, steer, a_x, nxt);
            if (P.flags & F_GYM_EGO) {                                   // E2E:281-282
                nxt[0] = (nxt[0] >= 0.0f) ? nxt[0] : 0.0f;
                nxt[5] = wrap_heading(nxt[5]);
            } else {
                nxt[0] = fminf(fmaxf(nxt[0], 0.0f), 35.0f);              // ego_predict, DM:390
            }
            int p = P.ref_idx ? P.ref_idx[rr] : P.path_index;
            const bool p_ok = (p >= 0) && (p < P.pv.n_paths);
            p = p_ok ? p : 0;
            const float2 *t_xy = s_xy + (size_t)p * P.pv.stride;
            int k0, k1;
            candidate_range(P.gv, p, (P.pv.N[p] + 1) & ~1, nxt[3], nxt[4], k0, k1);
            float best;
            int bi;
            scan_min(t_xy, k0, k1, nxt[3], nxt[4], best, bi);
            if (FUSED) {                                                  // state for the next step
                float t9[3] = {0.0f, 0.0f, 0.0f};
                if (p_ok)
                    tracking_from_index(t_xy, s_phi + (size_t)p * P.pv.stride, P.pv.L[p], P.pv.tail[p],
                                        P.task, bi, nxt[3], nxt[4], nxt[5], nxt[0], 0, t9);
#pragma unroll
                for (int i = 0; i < 6; ++i) e9n[i] = nxt[i];
                e9n[6] = t9[0]; e9n[7] = t9[1]; e9n[8] = t9[2];
            }
            if (valid && h == 1 && last_step) {
                float *q = P.obs_out + row * P.ld_out;
                float t9[3];
                if (p_ok) {
                    tracking_from_index(t_xy, s_phi + (size_t)p * P.pv.stride, P.pv.L[p],
                                        P.pv.tail[p], P.task, bi, nxt[3], nxt[4], nxt[5], nxt[0],
                                        0, t9);
                    if (P.n_future > 0)
                        tracking_from_index(t_xy, s_phi + (size_t)p * P.pv.stride, P.pv.L[p],
                                            P.pv.tail[p], P.task, bi, nxt[3], nxt[4], nxt[5], nxt[0],
                                            P.n_future, q + 6);
                } else {
                    t9[0] = t9[1] = t9[2] = 0.0f;
                    for (int i = 3; i < n_trk; ++i) q[6 + i] = 0.0f;     // DM:342-343
                }
                if (vec_out && veh_off == 9) {
                    q[0] = nxt[0];
                    *reinterpret_cast<float4 *>(q + 1) = make_float4(nxt[1], nxt[2], nxt[3], nxt[4]);
                    *reinterpret_cast<float4 *>(q + 5) = make_float4(nxt[5], t9[0], t9[1], t9[2]);
                } else {
#pragma unroll
                    for (int i = 0; i < 6; ++i) q[i] = nxt[i];
                    q[6] = t9[0]; q[7] = t9[1]; q[8] = t9[2];
                }
                if (P.act_scaled_out) {
                    P.act_scaled_out[2 * row] = steer;
                    P.act_scaled_out[2 * row + 1] = a_x;
                }
            }
        }

        // ---------------- vehicle phase ----------------
        float v2v_tr = 0.f, v2v_re = 0.f;        // this lane's half of the sums
        unsigned qa = q_lane;
        auto flush = [&]() {                      // finish this lane's queued pairs, in order
            const int cnt = (int)(qa - q_lane) >> 7;
#pragma unroll 1
            for (int i = 0; i < cnt; ++i) {
                // every queued dd is < 12.25, so sqrt(dd) - 3.5 < 0 holds (the gate is that test)
                const float d = __fsqrt_rn(lds_f32(q_lane + (unsigned)i * 128u));
                const float g35 = d - 3.5f, g25 = d - 2.5f;
                v2v_tr = v2v_tr + sq(g35);
                v2v_re = v2v_re + ((g25 < 0.0f) ? sq(g25) : 0.0f);
            }
            qa = q_lane;
        };
        if (FUSED && step == 0) {
            cp_async_wait<0>();
            __syncwarp();
        }
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int b = FUSED ? ch : (ch & 1);
            float *buf = scr.vbuf[b];
            if (!FUSED) {
                if (ch + 1 < n_chunks) stage(ch + 1, b ^ 1);
                else cp_async_commit();
                cp_async_wait<1>();
                __syncwarp();
            }
            float4 *slot = reinterpret_cast<float4 *>(buf + lane * (4 * VPL));
            const int j0 = h * H + ch * VPL;                // this lane's first vehicle of the chunk
            const int j_end = h ? P.V_in : min(H, P.V_in);  // end of this lane's half
            if (j0 + VPL <= j_end && (!NEXT || j0 + VPL <= P.V_out)) {
#pragma unroll
                for (int e = 0; e < VPL; e += 2) {          // two independent vehicles at a time
                    float4 v0 = slot[e ^ swz], v1 = slot[(e + 1) ^ swz];
                    v0 = vehicle_step<REW, NEXT, FAST>(v0, ec, P.turn_rs[j0 + e], P.turn_rr[j0 + e], P.turn_half[j0 + e], qa);
                    v1 = vehicle_step<REW, NEXT, FAST>(v1, ec, P.turn_rs[j0 + e + 1], P.turn_rr[j0 + e + 1], P.turn_half[j0 + e + 1], qa);
                    if (NEXT) { slot[e ^ swz] = v0; slot[(e + 1) ^ swz] = v1; }
                }
            } else {
                for (int e = 0; e < VPL; ++e) {
                    if (j0 + e < j_end) {
                        const float4 nv = vehicle_step<REW, NEXT, FAST>(slot[e ^ swz], ec, P.turn_rs[j0 + e], P.turn_rr[j0 + e], P.turn_half[j0 + e], qa);
                        if (NEXT && j0 + e < P.V_out) slot[e ^ swz] = nv;
                    }
                }
            }
            if (REW && __any_sync(0xffffffffu, (int)(qa - q_lane) > (QCAP - 4 * VPL) * 128)) flush();
            __syncwarp();
            if (NEXT && last_step) {
                float *dst = g_out + ch * (4 * VPL);
                const float *src = buf + p_soff;
                if (vec_out && rows_here == RPW && H + (ch + 1) * VPL <= P.V_out) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<float4 *>(dst + 4 * i * ld_out) =
                            *reinterpret_cast<const float4 *>(src + i * PIECE_STRIDE);
                } else if (p_veh + ch * VPL < (p_half ? P.V_out : min(H, P.V_out))) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (p_row + 4 * i < rows_here) {
                            float *d2 = dst + 4 * i * ld_out;
                            const float *s2 = src + i * PIECE_STRIDE;
                            if (vec_out) {
                                *reinterpret_cast<float4 *>(d2) = *reinterpret_cast<const float4 *>(s2);
                            } else {
                                d2[0] = s2[0]; d2[1] = s2[1]; d2[2] = s2[2]; d2[3] = s2[3];
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        cp_async_wait<0>();

        if (REW) {
            flush();
            // (first half) + (second half); the reward lane (h = 0, or both when !NEXT) writes
            const float tr_o = __shfl_xor_sync(0xffffffffu, v2v_tr, 1);
            const float re_o = __shfl_xor_sync(0xffffffffu, v2v_re, 1);
            if (valid && h == 0) {
                const float tr = v2v_tr + tr_o, re = v2v_re + re_o;
                float *o5 = P.out5 + (FUSED ? (int64_t)step * 5 * P.B : 0);
                o5[row] = rewards;
                o5[P.B + row] = tr + v2r_tr;                              // DM:299
                o5[2 * P.B + row] = re + v2r_re;                          // DM:300
                o5[3 * P.B + row] = re;
                o5[4 * P.B + row] = v2r_re;
                if (P.dict16) {
                    float *d = P.dict16 + row;
                    const int64_t B = P.B;
                    d[12 * B] = tr; d[13 * B] = v2r_tr; d[14 * B] = re; d[15 * B] = v2r_re;
                }
            }
        }
        if (FUSED) {                              // next step's ego + tracking columns: from the dynamics lane
#pragma unroll
            for (int i = 0; i < 9; ++i) e9[i] = __shfl_sync(0xffffffffu, e9n[i], lane | 1);
        }
        }                                         // step loop
    }
}

#include "ce2e_step_pair.cuh"

GridView make_grid_view(const ce2e_paths *p) {
    GridView g;
    memset(&g, 0, sizeof(g));
    g.cells = p->cells;
    g.inv_h = (float)(1.0 / GRID_H);
    for (int i = 0; i < p->n_paths; ++i) {
        g.off[i] = p->cell_off[i];
        g.nx[i] = p->grid[i].nx;
        g.ny[i] = p->grid[i].ny;
        g.x0[i] = p->grid[i].x0;
        g.y0[i] = p->grid[i].y0;
    }
    return g;
}

constexpr int CE2E_ERR_NOTMA = -1000;      // internal: fall back to k_model_step

// The warp-pair TMA kernel serves the fused rollout_out step on rows whose vehicle block is 16 B
// aligned in both buffers (what padded_rows / any ld % 4 == 0 layout with an aligned block gives).
bool pair_kernel_applies(const StepParams &P, const DeviceInfo *di) {
    const int need = F_REWARD | F_NEXT | F_VEC_IN | F_VEC_OUT;
    return g_use_tma && P.horizon == 0 && (P.flags & need) == need && P.V_in == P.V_out && P.V_in >= 1 &&
           P.B >= PAIR_ROWS && P.B < ((int64_t)1 << 31) - PAIR_ROWS &&
           (int)pair_smem_bytes(P.pv.n_paths, P.pv.stride) <= di->max_smem_optin;
}

int launch_model_step_pair(const StepParams &P, const DeviceInfo *di, cudaStream_t st) {
    PairParams PP;
    memset(&PP, 0, sizeof(PP));
    PP.S = P;
    const int veh_off = 6 + 3 * (P.n_future + 1), H0 = (P.V_in + 1) >> 1;
    for (int h = 0; h < 2; ++h) {
        const int n_veh = h ? P.V_in - H0 : H0;
        if (n_veh == 0) continue;
        const int off = veh_off + (h ? 4 * H0 : 0);
        if (!vehicle_tensor_map(P.obs_in + off, P.ld_in, P.B, n_veh, &PP.tm_in[h]) ||
            !vehicle_tensor_map(P.obs_out + off, P.ld_out, P.B, n_veh, &PP.tm_out[h]))
            return CE2E_ERR_NOTMA;
    }
    // ego + tracking columns as one more box per tile when they are the 9 floats right in front of the
    // vehicle block (n = 0); the store also rewrites the 28 B in front of each row, which must then be
    // padding of the previous row (ld - D >= 7, what padded_rows gives).  The maps start at row 1.
    if (veh_off == 9 && g_ego_tma) {
        if (g_ego_tma & 1)
            PP.ego_tma_in = vehicle_tensor_map(P.obs_in + P.ld_in - 7, P.ld_in, P.B - 1, 4, &PP.tm_ego_in, true);
        if ((g_ego_tma & 2) && P.ld_out - (9 + 4 * P.V_out) >= 7)
            PP.ego_tma_out = vehicle_tensor_map(P.obs_out + P.ld_out - 7, P.ld_out, P.B - 1, 4, &PP.tm_ego_out, true);
    }
    const bool fast = g_fast_trig;
    const int64_t n_tiles = (P.B + PAIR_ROWS - 1) / PAIR_ROWS;
    const int64_t max_blocks = PAIR_BLOCKS_PER_SM * (int64_t)di->sms;
    // Up to two tiles per pair (B <= 2 * 32 * 14 * SMs = 132608 rows on a B200): the reward warp runs its
    // vehicle chunks under the dynamics warp's latency chain (f_xu -> candidate cell -> scan -> tracking),
    // which measured 3-5 % faster than splitting that chain evenly (15.16 vs 15.65 us at B = 65536, 30.3 vs
    // 31.8 us at 131072).  Many tiles per pair: the even split (BAL) keeps both warps of a pair in step and
    // measured 6 % faster (117.0 vs 124.6 us at B = 524288).  CE2E_PAIR_BAL=0/1 forces.
    static const int env_bal = getenv("CE2E_PAIR_BAL") ? atoi(getenv("CE2E_PAIR_BAL")) : -1;
    const int force_bal = g_use_tma >= 2 ? g_use_tma - 2 : env_bal;
    // few vehicles: the reward warp streams both halves (PairParams::solo); measured break-even near V = 12
    static const int env_solo = getenv("CE2E_PAIR_SOLO_V") ? atoi(getenv("CE2E_PAIR_SOLO_V")) : 12;
    PP.solo = P.V_in <= env_solo;
    const bool bal = PP.solo ? false : (force_bal >= 0 ? force_bal != 0 : n_tiles > 2 * max_blocks * PAIR_PAIRS);
    void (*kern)(const PairParams);
    if (PP.solo) kern = fast ? k_model_step_pair<true, false, true> : k_model_step_pair<false, false, true>;
    else if (bal) kern = fast ? k_model_step_pair<true, true, false> : k_model_step_pair<false, true, false>;
    else kern = fast ? k_model_step_pair<true, false, false> : k_model_step_pair<false, false, false>;
    static std::atomic<bool> smem_set[64][6];
    int dev = 0;
    CE2E_CUDA(cudaGetDevice(&dev));
    std::atomic<bool> &set = smem_set[dev & 63][(fast ? 3 : 0) + (PP.solo ? 2 : (bal ? 1 : 0))];
    if (!set.load(std::memory_order_acquire)) {
        CE2E_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di->max_smem_optin));
        set.store(true, std::memory_order_release);
    }
    const int64_t blocks = n_tiles < max_blocks ? n_tiles : max_blocks;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(PAIR_WARPS * 32);
    cfg.dynamicSmemBytes = pair_smem_bytes(P.pv.n_paths, P.pv.stride);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, PP);
    ++g_launches;
    g_last_step_kernel = 2;
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CE2E_ERR_CUDA, "k_model_step_pair launch: %s", cudaGetErrorString(e));
    }
    return CE2E_OK;
}

int launch_model_step(StepParams &P, cudaStream_t st) {
    DeviceInfo *di;
    int rc = device_info(&di);
    if (rc) return rc;
    if (P.ld_in > (1 << 24) || P.ld_out > (1 << 24)) return fail(CE2E_ERR_SHAPE, "row stride too large");
    const bool fused = P.horizon > 0;
    if (pair_kernel_applies(P, di)) {
        rc = launch_model_step_pair(P, di, st);
        if (rc != CE2E_ERR_NOTMA) return rc;       // no tensor-map encoder in this driver: k_model_step below
    }
    const int64_t n_tiles = (P.B + RPW - 1) / RPW;
    size_t smem = (size_t)P.pv.n_paths * P.pv.stride * 12 + 32 +
                  STEP_WARPS * (fused ? sizeof(WarpScratchT<FUSED_MAX_CHUNKS>) : sizeof(WarpScratch));
    if ((int)smem > di->max_smem_optin)
        return fail(CE2E_ERR_SHAPE, "path tables need %zu B of shared memory (max %d)", smem,
                    di->max_smem_optin);
    void (*kern)(const StepParams) = nullptr;
    const bool rew = P.flags & F_REWARD, next = P.flags & F_NEXT;
    const bool fast = g_fast_trig && rew && next && !fused;
    if (fused) kern = k_model_step<true, true, false, true>;
    else if (fast) kern = k_model_step<true, true, true>;
    else if (rew && next) kern = k_model_step<true, true>;
    else if (rew) kern = k_model_step<true, false>;
    else kern = k_model_step<false, true>;
    // the opt-in shared-memory size is a process-wide, per-device attribute of each kernel
    // instantiation: raise it once to the device maximum (never to this call's size -- another host
    // thread launching with larger path tables must not find it lowered)
    static std::atomic<bool> smem_set[64][5];
    int dev = 0;
    CE2E_CUDA(cudaGetDevice(&dev));
    std::atomic<bool> &set = smem_set[dev & 63][fused ? 4 : fast ? 3 : rew && next ? 0 : rew ? 1 : 2];
    if (!set.load(std::memory_order_acquire)) {
        CE2E_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, di->max_smem_optin));
        set.store(true, std::memory_order_release);
    }
    // two persistent blocks per SM (one in the horizon-fused mode); never more blocks than tiles
    const int64_t max_blocks = (fused ? 1 : 2) * (int64_t)di->sms;
    const int64_t blocks = n_tiles < max_blocks ? n_tiles : max_blocks;
    // programmatic dependent launch: the next step's blocks may start (and stage their tables)
    // while this grid drains; they wait in griddepcontrol.wait before touching observations
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(STEP_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, P);
    ++g_launches;
    g_last_step_kernel = 1;
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CE2E_ERR_CUDA, "k_model_step launch: %s", cudaGetErrorString(e));
    }
    return CE2E_OK;
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
    sincos_cw(phi, s, c);
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
__global__ void k_tracking(const __grid_constant__ PathView pv, const __grid_constant__ GridView gv,
                           int task, int path_index, const int32_t *__restrict__ ref_idx,
                           const float *__restrict__ xs, const float *__restrict__ ys,
                           const float *__restrict__ phis, const float *__restrict__ vs, int n_future,
                           float *__restrict__ out, int64_t ld_out, int64_t B) {
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
    int bi, k0, k1;
    const float2 *t_xy = pv.xy + (size_t)p * pv.stride;
    candidate_range(gv, p, (pv.N[p] + 1) & ~1, x, y, k0, k1);
    scan_min(t_xy, k0, k1, x, y, best, bi);
    tracking_from_index(t_xy, pv.phi + (size_t)p * pv.stride, pv.L[p], pv.tail[p], task, bi, x, y,
                        phis[i], vs[i], n_future, q);
}

// find_closest_point, ratio 10, through the candidate grid (the path the fused step takes).
__global__ void k_closest10(const __grid_constant__ PathView pv, const __grid_constant__ GridView gv, int p,
                            const float *__restrict__ xs, const float *__restrict__ ys,
                            int64_t *__restrict__ idx_out, float *__restrict__ pts_out, int64_t B) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float x = xs[i], y = ys[i];
    float best;
    int bi, k0, k1;
    const float2 *t_xy = pv.xy + (size_t)p * pv.stride;
    candidate_range(gv, p, (pv.N[p] + 1) & ~1, x, y, k0, k1);
    scan_min(t_xy, k0, k1, x, y, best, bi);
    if (idx_out) idx_out[i] = 10 * (int64_t)bi;
    if (pts_out) {
        pts_out[i] = t_xy[bi].x;
        pts_out[B + i] = t_xy[bi].y;
        pts_out[2 * B + i] = pv.phi[(size_t)p * pv.stride + bi];
    }
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
    sincos_cw(th, s, c);
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
    sincos_cw(deg2rad(o[5]), s, c);
    const Circles e0 = circle_centres(o[3], o[4], s, c);
    sincos_cw(deg2rad(n[5]), s, c);
    const Circles e1 = circle_centres(n[3], n[4], s, c);
    float acc = 0.f;
    for (int j = 0; j < V; ++j) {
        const float *v = o + veh_off + 4 * j, *w = n + veh_off + 4 * j;
        const float ego2veh = __fsqrt_rn(sq(o[3] - v[0]) + sq(o[4] - v[1]));
        sincos_cw(deg2rad(v[3]), s, c);
        const Circles v0 = circle_centres(v[0], v[1], s, c);
        sincos_cw(deg2rad(w[3]), s, c);
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

// ------------------------------------------------------------------------------------------
// Vehicle stream of the small tiled kernels (k_env_done, k_model_step_bwd): k_model_step's
// decomposition without its persistent loop.  A warp owns RPW = 16 consecutive rows, two lanes per
// row; lane h visits half h of the row's vehicle list.  The 16 B vehicle records travel through
// shared memory in chunks of VPL = 4 per lane with coalesced cp.async copies (piece k of a lane =
// row (lane / 8) + 4 k of the tile, vehicle (lane % 4) of half (lane / 4) % 2; it lands in the slot
// of lane 2 * row + half at float4 index vehicle ^ (row % 4)), double buffered.  Needs a 16 B
// aligned vehicle block and ld % 4 == 0.
// ------------------------------------------------------------------------------------------
#ifndef CE2E_TILED_WARPS
#define CE2E_TILED_WARPS 4          // one partial wave at B = 65 536: 4-warp blocks spread evenly over the SMs (8: 19.7 vs 17.2 us)
#endif
constexpr int TILED_WARPS = CE2E_TILED_WARPS;
#ifndef CE2E_DONE_WARPS
#define CE2E_DONE_WARPS 2
#endif
// k_env_done is one partial wave (4096 warps at B = 65 536): small blocks spread it evenly over the SMs
constexpr int DONE_WARPS = CE2E_DONE_WARPS;
struct VehicleStream {
    float *buf;                // this warp's 2 x (32 * 4 * VPL) floats
    const float *g_in;
    int64_t ld;
    int V, H, n_chunks, rows_here, lane, p_row, p_half, p_veh, p_soff;

    __device__ __forceinline__ void init(float *warp_buf, const float *obs, int64_t ld_, int veh_off, int V_,
                                         int64_t tile, int64_t B, int lane_) {
        buf = warp_buf; ld = ld_; V = V_; lane = lane_;
        H = (V + 1) >> 1;
        n_chunks = (H + VPL - 1) / VPL;
        rows_here = (int)min((int64_t)RPW, B - tile * RPW);
        p_row = lane >> 3; p_half = (lane >> 2) & 1;
        const int p_e = lane & 3;
        p_veh = p_half * H + p_e;
        p_soff = (2 * p_row + p_half) * (4 * VPL) + ((p_e ^ (p_row & 3)) << 2);
        g_in = obs + tile * RPW * ld + (p_row * ld + veh_off + 4 * p_veh);
    }
    __device__ __forceinline__ void stage(int ch, int b) const {
        constexpr int PIECE_STRIDE = 8 * 4 * VPL;
        if (p_veh + ch * VPL < (p_half ? V : min(H, V))) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (p_row + 4 * k < rows_here)
                    cp_async16((unsigned)__cvta_generic_to_shared(buf + b * (32 * 4 * VPL) + p_soff + k * PIECE_STRIDE),
                               g_in + ch * (4 * VPL) + 4 * k * ld);
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void begin() const {
        if (n_chunks > 0) stage(0, 0);
    }
    // f(float4 vehicle) for every vehicle of this lane's half, in list order; begin() came first
    template <class F>
    __device__ __forceinline__ void run(F &&f) const {
        const int h = lane & 1, swz = (lane >> 1) & 3;
        const int j_end = h ? V : min(H, V);
        for (int ch = 0; ch < n_chunks; ++ch) {
            if (ch + 1 < n_chunks) stage(ch + 1, (ch + 1) & 1);
            else cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            const float4 *slot = reinterpret_cast<const float4 *>(buf + (ch & 1) * (32 * 4 * VPL) + lane * (4 * VPL));
            const int j0 = h * H + ch * VPL;
#pragma unroll
            for (int e = 0; e < VPL; ++e)
                if (j0 + e < j_end) f(slot[e ^ swz]);
            __syncwarp();
        }
        cp_async_wait<0>();
    }
    // f(float4 vehicle, bool valid) for every slot of every chunk, the whole warp converged in every call
    // (valid = the slot holds a vehicle of this lane's half), so f may use warp collectives
    template <class F>
    __device__ __forceinline__ void run_all(F &&f) const {
        const int h = lane & 1, swz = (lane >> 1) & 3;
        const int j_end = h ? V : min(H, V);
        for (int ch = 0; ch < n_chunks; ++ch) {
            if (ch + 1 < n_chunks) stage(ch + 1, (ch + 1) & 1);
            else cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            const float4 *slot = reinterpret_cast<const float4 *>(buf + (ch & 1) * (32 * 4 * VPL) + lane * (4 * VPL));
            const int j0 = h * H + ch * VPL;
#pragma unroll
            for (int e = 0; e < VPL; ++e) f(slot[e ^ swz], j0 + e < j_end);
            __syncwarp();
        }
        cp_async_wait<0>();
    }
};

// ------------------------------------------------------------------------------------------
// on-device reset of the batched environment (E2E:99-127, E2E:472-499)
// ------------------------------------------------------------------------------------------
// CrossroadEnd2end.reset -> _reset_init_state for every row whose done code is non-zero (or for all
// rows when `done` is NULL): a random waypoint index int(u * span) + 700 on the row's path (span =
// 1400 / 1700 / 920, E2E:473-478), ego = (v_x = 8 u', 0, 0, x, y, phi of that waypoint) (E2E:480-499),
// tracking columns = tracking_error_vector of that pose (E2E:293-297).  The reference then asks SUMO
// for the surrounding traffic (traffic.py:151-195, out of scope); here the V vehicle slots are drawn
// from the synthetic traffic distribution of SURVEY 8d (10 % within +-8 m of the ego, the rest uniform
// over the +-65 m map, pushed 30 m away if closer than 6 m; v ~ U(0, 8); heading on a compass
// direction + ~N(0, 10 deg) as the scaled sum of four uniforms).  All draws are Philox4x32-10 outputs
// keyed by `seed` at counter (row, episode[row], draw block); episode[row] is incremented.  One thread
// per row; rows that are not done return at once.
struct ResetParams {
    PathView pv;
    GridView gv;
    const float *full[CE2E_MAX_PATHS];
    int task, fixed_path, V, n_future, span;
    uint32_t seed_lo, seed_hi;
};

// One vehicle slot of a fresh episode (see k_env_reset): draws of Philox blocks 1 + 2 j and 2 + 2 j.
__device__ __forceinline__ float4 reset_vehicle(uint32_t i_lo, uint32_t i_hi, uint32_t ep, int j, uint32_t k0, uint32_t k1,
                                                float x, float y) {
    const Philox4 a = philox4x32_10(i_lo, i_hi, ep, 1u + 2u * (uint32_t)j, k0, k1);
    const Philox4 b = philox4x32_10(i_lo, i_hi, ep, 2u + 2u * (uint32_t)j, k0, k1);
    const bool near = (a.v[0] & 0xffffu) < 6554u;
    const int quad = (int)((a.v[0] >> 16) & 3u);
    const float u1 = u01_24(a.v[1]), u2 = u01_24(a.v[2]), u3 = u01_24(a.v[3]);
    float vx = near ? x + (u1 * 16.0f - 8.0f) : u1 * 130.0f - 65.0f;
    const float vy = near ? y + (u2 * 16.0f - 8.0f) : u2 * 130.0f - 65.0f;
    const float dd = sq(vx - x) + sq(vy - y);
    vx = (dd < 36.0f) ? vx + 30.0f : vx;
    const float s4 = ((u01_24(b.v[0]) + u01_24(b.v[1])) + u01_24(b.v[2])) + u01_24(b.v[3]);
    const float base = quad == 0 ? 0.0f : (quad == 1 ? 90.0f : (quad == 2 ? 180.0f : -90.0f));
    float vphi = base + 10.0f * ((s4 - 2.0f) * 1.73205077648162842f);
    vphi = (vphi > 180.0f) ? vphi - 360.0f : vphi;
    vphi = (vphi <= -180.0f) ? vphi + 360.0f : vphi;
    return make_float4(vx, vy, 8.0f * u3, vphi);
}

// Ego + tracking columns, path index and red-light flag of row i's next episode (one thread); returns the
// ego position through x, y and the episode number just consumed through ep.
__device__ __forceinline__ void reset_ego(const ResetParams &R, int64_t i, int32_t *episode, float *obs, int64_t ld,
                                          int32_t *ref_idx, int8_t *virtual_red, uint32_t &ep, float &x, float &y) {
    ep = (uint32_t)episode[i];
    episode[i] = (int32_t)(ep + 1u);
    const Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), ep, 0u, R.seed_lo, R.seed_hi);
    const int p = R.fixed_path >= 0 ? R.fixed_path : (int)(((uint64_t)r.v[0] * (uint32_t)R.pv.n_paths) >> 32);
    const int L = R.pv.L[p];
    int idx = (int)(((uint64_t)r.v[1] * (uint32_t)R.span) >> 32) + 700;         // E2E:473-478
    idx = idx < L ? idx : L - 1;                                                 // indexs2points clamp, DM:727-728
    const float *tab = R.full[p];
    x = tab[idx];
    y = tab[L + idx];
    const float phi = tab[2 * (size_t)L + idx];
    const float v = CE2E_EXP_V * u01_24(r.v[2]);                                 // E2E:482
    float *o = obs + i * ld;
    o[0] = v; o[1] = 0.0f; o[2] = 0.0f; o[3] = x; o[4] = y; o[5] = phi;
    int k0, k1, bi;
    float best;
    const float2 *t_xy = R.pv.xy + (size_t)p * R.pv.stride;
    candidate_range(R.gv, p, (R.pv.N[p] + 1) & ~1, x, y, k0, k1);
    scan_min(t_xy, k0, k1, x, y, best, bi);
    tracking_from_index(t_xy, R.pv.phi + (size_t)p * R.pv.stride, L, R.pv.tail[p], R.task, bi, x, y, phi, v,
                        R.n_future, o + 6);
    ref_idx[i] = p;
    if (virtual_red) virtual_red[i] = (int8_t)(u01_24(r.v[3]) > 0.9f);             // E2E:120-124
}

// The vehicle slots of the rows flagged in `pending` (bit = lane that holds the row's ep / x / y; row =
// row0 + (lane >> lane_shift)), generated by the whole warp: lane j takes slots j, j + 32, ...
__device__ __forceinline__ void reset_vehicles(const ResetParams &R, unsigned pending, int lane_shift, int64_t row0,
                                               uint32_t ep, float x, float y, float *obs, int64_t ld, int lane) {
    const int veh_off = 6 + 3 * (R.n_future + 1);
    while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const int64_t ri = row0 + (src >> lane_shift);
        const uint32_t rep = __shfl_sync(0xffffffffu, ep, src);
        const float rx = __shfl_sync(0xffffffffu, x, src), ry = __shfl_sync(0xffffffffu, y, src);
        float *veh = obs + ri * ld + veh_off;
        for (int j = lane; j < R.V; j += 32) {
            const float4 w = reset_vehicle((uint32_t)ri, (uint32_t)((uint64_t)ri >> 32), rep, j, R.seed_lo, R.seed_hi, rx, ry);
            if ((reinterpret_cast<uintptr_t>(veh) & 15u) == 0) {
                reinterpret_cast<float4 *>(veh)[j] = w;
            } else {
                veh[4 * j] = w.x; veh[4 * j + 1] = w.y; veh[4 * j + 2] = w.z; veh[4 * j + 3] = w.w;
            }
        }
    }
}

// A warp owns 32 consecutive rows.  Phase 1, one lane per row: the rows to reset draw their ego state, read
// the waypoint, project it (tracking columns) and write columns 0 .. veh_off - 1 -- a chain of dependent
// global loads, so all rows of the warp take it at the same time.  Phase 2, the whole warp per reset row
// (found with one ballot; ego position and episode number broadcast by shuffle): lane j (j, j + 32, ...)
// generates vehicle slot j.  (With everything in one thread per row a launch was as slow as ~65 Philox
// blocks in sequence for V = 32.)
__global__ void __launch_bounds__(256)
k_env_reset(const __grid_constant__ ResetParams R, int32_t *__restrict__ episode,
            const int8_t *__restrict__ done, float *__restrict__ obs, int64_t ld,
            int32_t *__restrict__ ref_idx, int8_t *__restrict__ virtual_red, uint8_t *__restrict__ done_flag, int64_t B) {
    const int lane = threadIdx.x & 31;
    const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (row0 >= B) return;
    const int64_t i = row0 + lane;
    const bool todo = i < B && (!done || done[i] != 0);
    if (done_flag && i < B) done_flag[i] = (uint8_t)todo;
    const unsigned pending = __ballot_sync(0xffffffffu, todo);
    if (!pending) return;
    uint32_t ep = 0;
    float x = 0.f, y = 0.f;
    if (todo) reset_ego(R, i, episode, obs, ld, ref_idx, virtual_red, ep, x, y);
    reset_vehicles(R, pending, 0, row0, ep, x, y, obs, ld, lane);
}

// CrossroadEnd2end._judge_done (E2E:200-256) on the observation AFTER a step, one thread per row.
// Order of the checks as in the reference: collision (Traffic.collision_check, traffic.py:263-295,
// with every surrounding vehicle taken as L x W = 4.8 x 2.0 like the ego), road constraint on the
// four body corners (E2E:171-177, EU:73-104), |delta_y| > 15 (E2E:223-225), yaw-rate bound
// r_bound = miu_r * g / (|v_x| + 1e-8) with miu_r of the step just taken (E2E:167, 231-242),
// red light (E2E:244-245; v_light is an input), goal box (E2E:247-256).
// code: 0 not_done_yet, 1 collision, 2 break_road_constrain, 3 deviate_too_much,
//       4 break_stability, 5 break_red_light, 6 good_done.
__device__ __forceinline__ bool feasible_point(int task, float px, float py) {
    const float road = CE2E_LW3;
    if (px > -CE2E_HALF && px < CE2E_HALF && py > -CE2E_HALF && py < CE2E_HALF) return true;
    if (task == 0) return (px > 0.f && px < CE2E_LW && py <= -CE2E_HALF) || (py > 0.f && py < road && px < -CE2E_HALF);
    if (task == 1) return (px > CE2E_LW && px < CE2E_LW2 && py <= -CE2E_HALF) || (px > 0.f && px < road && py >= CE2E_HALF);
    return (px > CE2E_LW2 && px < road && py <= -CE2E_HALF) || (py > -road && py < 0.f && px > CE2E_HALF);
}

// TILED = false: one thread per row, scalar loads (any alignment).  TILED = true: VehicleStream,
// the collision test branch free (a quarter of the synthetic vehicles sit inside the 10 m gate, so a
// branch would run for nearly every vehicle with a few live lanes), the two halves OR-ed by shuffle.
// RESET (tiled kernel only): the rows that are done start their next episode in the same launch
// (reset_ego / reset_vehicles above = ce2e_env_reset), and done_flag receives code != 0 as 0 / 1 bytes.
struct DoneReset {
    ResetParams R;
    int32_t *episode, *ref_idx;
    int8_t *virtual_red;
    uint8_t *done_flag;
    float *obs_w;                            // the same rows as `obs`, writable
};
template <bool TILED, bool RESET = false>
__global__ void __launch_bounds__(TILED ? DONE_WARPS * 32 : 128)
k_env_done(const __grid_constant__ DynConsts K, int task, const float *__restrict__ obs,
           int64_t ld, const float *__restrict__ act_scaled, int V, int veh_off,
           int v_light, int8_t *__restrict__ done, int64_t B, const __grid_constant__ DoneReset X) {
    __shared__ __align__(16) float s_veh[TILED ? DONE_WARPS : 1][TILED ? 2 * 32 * 4 * VPL : 4];
    __shared__ float4 s_gate[TILED ? DONE_WARPS : 1][TILED ? 64 : 1];      // vehicles inside the gate, per warp
    __shared__ float4 s_ego[TILED ? DONE_WARPS : 1][TILED ? RPW : 1];      // the rows' ego circle centres
    __shared__ int s_hit[TILED ? DONE_WARPS : 1][TILED ? RPW : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * DONE_WARPS + warp;
    int64_t i;
    bool owner = true, live = true;
    VehicleStream vs;
    if (TILED) {
        if (tile * RPW >= B) return;         // whole warp
        const int64_t row = tile * RPW + (lane >> 1);
        live = row < B;
        owner = live && (lane & 1) == 0;
        i = live ? row : B - 1;
        vs.init(s_veh[warp], obs, ld, veh_off, V, tile, B, lane);
        vs.begin();
    } else {
        i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= B) return;
    }
    const float *o = obs + i * ld;
    const float vx = o[0], r = o[2], x = o[3], y = o[4], phi = o[5], dy = o[6];
    float s, c;
    sincos_cw(deg2rad(phi), s, c);
    int code = 0;
    // collision: two-circle model, 10 m box gate, threshold ((w + w)/2 + 0.5)^2
    {
        const Circles e = circle_centres(x, y, s, c);
        const float thr = 6.25f;
        bool hit = false;
        if (TILED) {
            // Few vehicles are inside the 10 m gate, but nearly every warp has some lane with one: the lanes
            // only test the gate and push the vehicles that pass into a warp-shared ring; whenever it holds 32,
            // every lane takes one (heading sin / cos, circle centres, the four distances).  The result is an OR
            // over the row's vehicles, so the order does not matter.
            const int rloc = lane >> 1;
            if ((lane & 1) == 0) {
                s_ego[warp][rloc] = make_float4(e.fx, e.fy, e.rx, e.ry);
                s_hit[warp][rloc] = 0;
            }
            __syncwarp();
            int head = 0, count = 0;                                       // warp-uniform
            auto take = [&](int k) {
                const float4 q = s_gate[warp][k & 63];
                const int r = __float_as_int(q.w);
                const float4 eg = s_ego[warp][r];
                float ws, wc;
                sincos_cw(deg2rad(q.z), ws, wc);
                const Circles w = circle_centres(q.x, q.y, ws, wc);
                const bool in = (sq(eg.x - w.fx) + sq(eg.y - w.fy) < thr) | (sq(eg.x - w.rx) + sq(eg.y - w.ry) < thr) |
                                (sq(eg.z - w.rx) + sq(eg.w - w.ry) < thr) | (sq(eg.z - w.fx) + sq(eg.w - w.fy) < thr);
                if (in) s_hit[warp][r] = 1;
            };
            vs.run_all([&](float4 v, bool valid) {
                const bool gate = valid && live && fabsf(v.x - x) < 10.0f && fabsf(v.y - y) < 10.0f;
                const unsigned m = __ballot_sync(0xffffffffu, gate);
                if (gate)
                    s_gate[warp][(head + count + __popc(m & ((1u << lane) - 1u))) & 63] =
                        make_float4(v.x, v.y, v.w, __int_as_float(rloc));
                count += __popc(m);
                if (count >= 32) {
                    __syncwarp();
                    take(head + lane);
                    head += 32; count -= 32;
                    __syncwarp();
                }
            });
            __syncwarp();
            if (lane < count) take(head + lane);
            __syncwarp();
            hit = s_hit[warp][rloc] != 0;
            if (!RESET && !owner) return;
        } else {
            for (int j = 0; j < V; ++j) {
                const float *v = o + veh_off + 4 * j;
                if (fabsf(v[0] - x) < 10.0f && fabsf(v[1] - y) < 10.0f) {
                    float ws, wc;
                    sincos_cw(deg2rad(v[3]), ws, wc);
                    const Circles w = circle_centres(v[0], v[1], ws, wc);
                    hit = hit || (sq(e.fx - w.fx) + sq(e.fy - w.fy) < thr) || (sq(e.fx - w.rx) + sq(e.fy - w.ry) < thr) ||
                          (sq(e.rx - w.rx) + sq(e.ry - w.ry) < thr) || (sq(e.rx - w.fx) + sq(e.ry - w.fy) < thr);
                }
            }
        }
        if (hit) code = 1;
    }
    if (code == 0) {                                     // four corners (+-l/2, +-w/2) in the world frame
        const float hl = 2.4f, hw = 1.0f;
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float lx = (k & 2) ? -hl : hl, ly = (k & 1) ? -hw : hw;
            const float px = (lx * c - ly * s) + x, py = (lx * s + ly * c) + y;
            ok = ok && feasible_point(task, px, py);
        }
        if (!ok) code = 2;
    }
    if (code == 0 && fabsf(dy) > 15.0f) code = 3;
    if (code == 0) {
        const float a_x = act_scaled[2 * i + 1];
        const float half_ma = (K.m * a_x) / 2.0f;
        const float F_xr = (a_x < 0.0f) ? half_ma : K.m * a_x;
        const float miu_r = sqrtf(sq(K.muFzr) - sq(F_xr)) / K.Fzr;      // DM:69
        const float r_bound = (miu_r * 9.81f) / (fabsf(vx) + 1e-8f);
        if (!(-r_bound < r && r < r_bound)) code = 4;
    }
    if (code == 0 && v_light != 0 && y > -CE2E_HALF && task != 2) code = 5;
    if (code == 0) {
        bool goal;
        if (task == 0) goal = x < -CE2E_HALF - 10.0f && y > 0.f && y < CE2E_LW3;
        else if (task == 2) goal = x > CE2E_HALF + 10.0f && y > -CE2E_LW3 && y < 0.f;
        else goal = y > CE2E_HALF + 10.0f && x > 0.f && x < CE2E_LW3;
        if (goal) code = 6;
    }
    if (!RESET) {
        done[i] = (int8_t)code;
        return;
    }
    if (owner) {
        done[i] = (int8_t)code;
        if (X.done_flag) X.done_flag[i] = (uint8_t)(code != 0);
    }
    // the warp's 16 rows have been read completely (vs.run is over): rows that are done restart here
    const bool todo = owner && code != 0;
    const unsigned pending = __ballot_sync(0xffffffffu, todo);
    if (!pending) return;
    uint32_t ep = 0;
    float rx = 0.f, ry = 0.f;
    if (todo) reset_ego(X.R, i, X.episode, X.obs_w, ld, X.ref_idx, X.virtual_red, ep, rx, ry);
    reset_vehicles(X.R, pending, 1, tile * RPW, ep, rx, ry, X.obs_w, ld, lane);
}

// ------------------------------------------------------------------------------------------
// backward of one rollout_out step
// ------------------------------------------------------------------------------------------
// Vector-Jacobian product of EnvironmentModel.rollout_out (DM:118-126) as TensorFlow's autodiff
// defines it in the reference: vehicle columns carry tf.stop_gradient (DM:195, DM:331, DM:402), the
// closest-waypoint index is an integer (tf.argmin / tf.gather: the reference point is a constant),
// tf.where passes the gradient of the selected branch, tf.clip_by_value passes it inside the
// closed interval.  One thread per row; the forward intermediates are recomputed.
//   inputs : obs_in, act_norm (the forward inputs), upstream gradients g_next [B, 9+3n] (d loss /
//            d next ego + next tracking columns) and g_out5 [5, B]
//   outputs: g_obs [B, 9+3n] (ego + tracking columns of obs_in; the future-preview tracking columns
//            only feed nothing and get 0), g_act [B, 2] (w.r.t. the NORMALISED actions)
__device__ __forceinline__ void road_terms_grad(int task, float px, float py, float w_tr, float w_re,
                                                float &gpx, float &gpy) {
    // d/d(px, py) of the four hinge^2 terms of road_terms(); w_tr / w_re weight the training / real sums
    if (task == 0) {
        const bool before = py < -CE2E_HALF, after = px < -CE2E_HALF;
        const float g1 = px - 1.0f, g2 = (CE2E_LW - px) - 1.0f, g3 = (CE2E_LW3 - py) - 1.0f, g4 = (py - 0.0f) - 1.0f;
        const float w = w_tr + w_re;
        if (before && px < 1.0f) gpx += w * 2.0f * g1;
        if (before && (CE2E_LW - px) < 1.0f) gpx -= w * 2.0f * g2;
        if (px < 0.0f && (CE2E_LW3 - py) < 1.0f) gpy -= w_tr * 2.0f * g3;
        if (after && (CE2E_LW3 - py) < 1.0f) gpy -= w_re * 2.0f * g3;
        if (after && (py - 0.0f) < 1.0f) gpy += w * 2.0f * g4;
    } else if (task == 1) {
        const bool before = py < -CE2E_HALF, after = py > CE2E_HALF;
        const float g1 = (px - CE2E_LW) - 1.0f, g2 = (CE2E_LW2 - px) - 1.0f, g3 = (CE2E_LW3 - px) - 1.0f, g4 = (px - 0.0f) - 1.0f;
        const float w = w_tr + w_re;
        if (before && (px - CE2E_LW) < 1.0f) gpx += w * 2.0f * g1;
        if (before && (CE2E_LW2 - px) < 1.0f) gpx -= w * 2.0f * g2;
        if (after && (CE2E_LW3 - px) < 1.0f) gpx -= w * 2.0f * g3;
        if (after && (px - 0.0f) < 1.0f) gpx += w * 2.0f * g4;
    } else {
        const bool before = py < -CE2E_HALF, after = px > CE2E_HALF;
        const float g1 = (px - CE2E_LW2) - 1.0f, g2 = (CE2E_LW3 - px) - 1.0f, g3 = (0.0f - py) - 1.0f, g4 = (py - (-CE2E_LW3)) - 1.0f;
        const float w = w_tr + w_re;
        if (before && (px - CE2E_LW2) < 1.0f) gpx += w * 2.0f * g1;
        if (before && (CE2E_LW3 - px) < 1.0f) gpx -= w * 2.0f * g2;
        if (after && (0.0f - py) < 1.0f) gpy -= w * 2.0f * g3;
        if (after && (py - (-CE2E_LW3)) < 1.0f) gpy += w * 2.0f * g4;
    }
}

__device__ __forceinline__ void pair_grad(float ex, float ey, float px, float py, float w_tr, float w_re,
                                          float &gx, float &gy) {
    // d/dE of hinge^2: 2 (d - thr) (E - P) / d inside the threshold, else 0.  Branch free (few lanes
    // of a warp are ever inside the gate).  1/d: the special-function unit's estimate plus one Newton
    // step (2 FMA), ~1 ulp -- the gradients are checked at 1e-4 against the reference's own autodiff.
    const float dx = ex - px, dy = ey - py;
    const float dd = dx * dx + dy * dy;
    float inv_d = rsqrtf(dd);
    inv_d = __fmaf_rn(inv_d, __fmaf_rn(-0.5f * dd * inv_d, inv_d, 0.5f), inv_d);
    const float d = dd * inv_d;
    const float k = (w_tr * fminf(d - 3.5f, 0.0f) + w_re * fminf(d - 2.5f, 0.0f)) * (2.0f * inv_d);
    const bool in = dd < 12.25f && dd > 0.0f;
    gx += in ? k * dx : 0.0f;
    gy += in ? k * dy : 0.0f;
}

// TILED = false: one thread per row, scalar loads (any alignment).
// TILED = true : VehicleStream -- a warp owns 16 rows, two lanes per row; lane h sums the collision
//                gradient over vehicle half h; then lane 0 takes the reward / road terms and lane 1
//                the next-observation terms.  Needs a 16 B aligned vehicle block, ld % 4 == 0.
template <bool TILED>
__global__ void __launch_bounds__(TILED ? TILED_WARPS * 32 : 128, TILED ? 4 : 1)
k_model_step_bwd(const __grid_constant__ PathView pv, const __grid_constant__ GridView gv,
                 const __grid_constant__ DynConsts K, int task, int path_index,
                 const int32_t *__restrict__ ref_idx, const float *__restrict__ obs,
                 int64_t ld, const float *__restrict__ act_norm, int V, int n_future,
                 const float *__restrict__ g_next, int64_t ld_gn,
                 const float *__restrict__ g_out5, float *__restrict__ g_obs, int64_t ld_go,
                 float *__restrict__ g_act, int64_t B) {
    __shared__ __align__(16) float s_veh[TILED ? TILED_WARPS : 1][TILED ? 2 * 32 * 4 * VPL : 4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = (int64_t)blockIdx.x * TILED_WARPS + warp;
    int64_t i;
    bool owner = true;                       // this thread writes the row's outputs
    if (TILED) {
        if (tile * RPW >= B) return;         // whole warp
        const int64_t row = tile * RPW + (lane >> 1);
        owner = row < B && (lane & 1) == 0;
        i = row < B ? row : B - 1;
    } else {
        i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= B) return;
    }
    const float *o = obs + i * ld;
    const int n_trk = 3 * (n_future + 1);
    VehicleStream vs;
    if (TILED) {
        vs.init(s_veh[warp], obs, ld, 6 + n_trk, V, tile, B, lane);
        vs.begin();                                   // in flight while the ego columns arrive
    }
    const float vx = o[0], vy = o[1], r = o[2], x = o[3], y = o[4], phi_deg = o[5];
    const float a0 = act_norm[2 * i], a1 = act_norm[2 * i + 1];
    float steer, a_x;
    action_transform(a0, a1, steer, a_x);
    const float m0 = (a0 >= -1.05f && a0 <= 1.05f) ? 0.4f : 0.0f;        // d steer / d a0
    const float m1 = (a1 >= -1.05f && a1 <= 1.05f) ? 2.25f : 0.0f;       // d a_x / d a1
    const float D2R = CE2E_PI32 / 180.0f;
    const float phi = deg2rad(phi_deg);
    float s, c;
    sincos_cw(phi, s, c);
    const float gR = g_out5[i], gPtr = g_out5[B + i], gPre = g_out5[2 * B + i], gV2V = g_out5[3 * B + i],
                gV2R = g_out5[4 * B + i];

    // gradients w.r.t. (vx, vy, r, x, y, phi_deg, steer, a_x) and the tracking columns
    float gvx = 0.f, gvy = 0.f, gr = 0.f, gx = 0.f, gy = 0.f, gphi = 0.f, gsteer = 0.f, gax = 0.f;
    // ---- rewards (DM:198-207, 297-298)
    const float g_dy = gR * (-(2.0f * CE2E_K.w_y) * o[6]);
    const float g_dphi = gR * (-(2.0f * CE2E_K.w_phi) * D2R * D2R * o[7]);
    const float g_dv = gR * (-(2.0f * CE2E_K.w_v) * o[8]);
    // TILED: lane h = 0 takes the reward + penalty terms, lane h = 1 the next-observation terms; the
    // partial gradients are added at the end
    const bool do_rew = !TILED || (lane & 1) == 0, do_next = !TILED || (lane & 1) == 1;
    if (do_rew) {
        gr += gR * (-(2.0f * CE2E_K.w_yaw) * r);
        gsteer += gR * (-(2.0f * CE2E_K.w_steer) * steer);
        gax += gR * (-(2.0f * CE2E_K.w_ax) * a_x);
    }
    // ---- collision and road penalties through the ego circle centres (DM:210-295)
    {
        const Circles ec = circle_centres(x, y, s, c);
        const float w_tr = gPtr, w_re_v = gPre + gV2V, w_re_r = gPre + gV2R;
        float gfx = 0.f, gfy = 0.f, grx = 0.f, gry = 0.f;
        const float *veh = o + 6 + n_trk;
        auto one_vehicle = [&](float vx_, float vy_, float vphi) {
            float vs, vc;
            sincos_cw(deg2rad(vphi), vs, vc);
            const Circles w = circle_centres(vx_, vy_, vs, vc);
            pair_grad(ec.fx, ec.fy, w.fx, w.fy, w_tr, w_re_v, gfx, gfy);
            pair_grad(ec.fx, ec.fy, w.rx, w.ry, w_tr, w_re_v, gfx, gfy);
            pair_grad(ec.rx, ec.ry, w.fx, w.fy, w_tr, w_re_v, grx, gry);
            pair_grad(ec.rx, ec.ry, w.rx, w.ry, w_tr, w_re_v, grx, gry);
        };
        if (TILED) {
            vs.run([&](float4 v) { one_vehicle(v.x, v.y, v.w); });
            // the two halves of the row
            gfx += __shfl_xor_sync(0xffffffffu, gfx, 1);
            gfy += __shfl_xor_sync(0xffffffffu, gfy, 1);
            grx += __shfl_xor_sync(0xffffffffu, grx, 1);
            gry += __shfl_xor_sync(0xffffffffu, gry, 1);
        } else {
            for (int j = 0; j < V; ++j) one_vehicle(veh[4 * j], veh[4 * j + 1], veh[4 * j + 3]);
        }
        if (do_rew) {
            road_terms_grad(task, ec.fx, ec.fy, w_tr, w_re_r, gfx, gfy);
            road_terms_grad(task, ec.rx, ec.ry, w_tr, w_re_r, grx, gry);
            gx += gfx + grx;
            gy += gfy + gry;
            // F = (x + l c, y + l s), R = (x - l c, y - l s);  d/dphi_deg = D2R * d/dtheta
            gphi += D2R * CE2E_LWS * ((-s) * gfx + c * gfy + s * grx - c * gry);
        }
    }
    if (do_next) {
        // ---- next observation (DM:322-353)
        const float *gn = g_next + i * ld_gn;
        float nxt[6];
        f_xu_next(K, vx, vy, r, x, y, phi, s, c, steer, a_x, nxt);
        const float vx_raw = nxt[0];
        nxt[0] = fminf(fmaxf(vx_raw, 0.0f), 35.0f);
        float g_n[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) g_n[k] = gn[k];
        int p = ref_idx ? ref_idx[i] : path_index;
        if (p >= 0 && p < pv.n_paths) {                                     // tracking' = T(x', y', phi', vx')
            // the closest waypoint is an integer index (tf.argmin): the reference point is a constant
            // of the derivative, so no waypoint scan is needed here
            const float ex = nxt[3], ey = nxt[4];
            float ddx = 0.f, ddy = 0.f;                                     // d(two2one)/d(ex, ey) = -d(delta)
            if (task == 0) {
                const float rho = __fsqrt_rn(sq(ex + CE2E_HALF) + sq(ey + CE2E_HALF));
                ddx = -(ex + CE2E_HALF) / rho; ddy = -(ey + CE2E_HALF) / rho;
                if (ey < -CE2E_HALF) { ddx = -1.0f; ddy = 0.0f; }
                if (ex < -CE2E_HALF) { ddx = 0.0f; ddy = -1.0f; }
            } else if (task == 1) {
                ddx = -1.0f;
            } else {
                const float rho = __fsqrt_rn(sq(ex - CE2E_HALF) + sq(ey + CE2E_HALF));
                ddx = (ex - CE2E_HALF) / rho; ddy = (ey + CE2E_HALF) / rho;
                if (ey < -CE2E_HALF) { ddx = -1.0f; ddy = 0.0f; }
                if (ex > CE2E_HALF) { ddx = 0.0f; ddy = 1.0f; }
            }
            g_n[3] += gn[6] * ddx;
            g_n[4] += gn[6] * ddy;
            g_n[5] += gn[7];
            g_n[0] += gn[8];
            for (int k = 0; k < n_future; ++k) {                             // (fx - x', fy - y', wrap(phi' - fphi))
                g_n[3] -= gn[9 + 3 * k];
                g_n[4] -= gn[10 + 3 * k];
                g_n[5] += gn[11 + 3 * k];
            }
        }
        if (!(vx_raw >= 0.0f && vx_raw <= 35.0f)) g_n[0] = 0.0f;            // clip_by_value (DM:390)
        // ---- back through f_xu (DM:73-81)
        {
            const float tau = K.tau;
            // vx' = vx + tau (a_x + vy r)
            gvx += g_n[0]; gvy += g_n[0] * tau * r; gr += g_n[0] * tau * vy; gax += g_n[0] * tau;
            // vy' = N1 / D1
            const float D1 = K.m * vx - K.Dv, vy1 = nxt[1];
            const float k1 = g_n[1] / D1;
            gvx += k1 * ((K.m * vy - K.tauCf * steer - 2.0f * K.taum * vx * r) - vy1 * K.m);
            gvy += k1 * (K.m * vx);
            gr += k1 * (K.tauK1 - K.taum * vx * vx);
            gsteer += k1 * (-K.tauCf * vx);
            // r' = N2 / D2
            const float D2 = K.Dr - K.Iz * vx, r1 = nxt[2];
            const float k2 = g_n[2] / D2;
            gvx += k2 * ((-K.Iz * r + K.tauaCf * steer) + r1 * K.Iz);
            gvy += k2 * (-K.tauK1);
            gr += k2 * (-K.Iz * vx);
            gsteer += k2 * (K.tauaCf * vx);
            // x' = x + tau (vx c - vy s), y' = y + tau (vx s + vy c)
            gx += g_n[3]; gy += g_n[4];
            gvx += tau * (g_n[3] * c + g_n[4] * s);
            gvy += tau * (-g_n[3] * s + g_n[4] * c);
            gphi += D2R * tau * (g_n[3] * (-vx * s - vy * c) + g_n[4] * (vx * c - vy * s));
            // phi' = phi + tau r 180/pi
            gphi += g_n[5];
            gr += g_n[5] * tau * (180.0f / CE2E_PI32);
        }
    }
    if (TILED) {
        gvx += __shfl_xor_sync(0xffffffffu, gvx, 1);
        gvy += __shfl_xor_sync(0xffffffffu, gvy, 1);
        gr += __shfl_xor_sync(0xffffffffu, gr, 1);
        gx += __shfl_xor_sync(0xffffffffu, gx, 1);
        gy += __shfl_xor_sync(0xffffffffu, gy, 1);
        gphi += __shfl_xor_sync(0xffffffffu, gphi, 1);
        gsteer += __shfl_xor_sync(0xffffffffu, gsteer, 1);
        gax += __shfl_xor_sync(0xffffffffu, gax, 1);
        if (!owner) return;
    }
    float *go = g_obs + i * ld_go;
    go[0] = gvx; go[1] = gvy; go[2] = gr; go[3] = gx; go[4] = gy; go[5] = gphi;
    go[6] = g_dy; go[7] = g_dphi; go[8] = g_dv;
    for (int k = 9; k < 6 + n_trk; ++k) go[k] = 0.0f;
    g_act[2 * i] = gsteer * m0;
    g_act[2 * i + 1] = gax * m1;
}

// ------------------------------------------------------------------------------------------
// interested-vehicle selection
// ------------------------------------------------------------------------------------------
// CrossroadEnd2end._construct_veh_vector_short (E2E:340-464) for B scenes: from the N vehicles
// around each ego keep, per route class of VEHICLE_MODE_DICT[task] (EU:21-23), the vehicles that
// pass the class's range filter (E2E:393-411), order them by the class's sort key (E2E:414-428,
// Python's stable sort: equal keys keep their list order) and emit the first `num`, padding with
// the class's far-away fill vehicle (E2E:440-447).  One thread per (scene, class).
// Route classes (the order of the reference's local lists, E2E:354): 0 dl, 1 du, 2 dr, 3 rd, 4 rl,
// 5 ru, 6 ur, 7 ud, 8 ul, 9 lu, 10 lr, 11 ld; anything else is ignored.
struct SelectSpec {
    int n_modes;
    int cls[5];        // route class of each selected mode, in VEHICLE_MODE_DICT[task] order
    int num[5];        // vehicles kept per mode
    int slot0[5];      // first output slot of the mode
};

__device__ __forceinline__ bool veh_in_range(int cls, int task, float x, float y, float ex, float ey) {
    switch (cls) {
        case 0: return x > -35.0f && y > ey - 2.0f;                                          // dl
        case 1: return ey - 2.0f < y && y < 35.0f && x < ex + 5.0f;                          // du
        case 2: return x < 35.0f && y > ey;                                                  // dr
        case 5: return x < 35.0f && y < 35.0f;                                               // ru
        case 6: return task == 1 ? (x < ex + 7.0f && ey < y && y < 35.0f)                    // ur
                                 : (task == 2 ? (x < 35.0f && y < 25.0f) : true);
        case 7: return fmaxf(ey - 2.0f, -25.0f) < y && y < 25.0f && ex > x;                  // ud
        case 8: return -35.0f < x && x < ex && y < 25.0f;                                    // ul
        case 10: return -35.0f < x && x < 35.0f;                                             // lr
        default: return true;
    }
}
// ascending lexicographic key (k1, k2) equivalent to the reference's sorted(...) call
__device__ __forceinline__ void veh_sort_key(int cls, int task, float x, float y, float &k1, float &k2) {
    switch (cls) {
        case 0: k1 = y; k2 = -x; break;                       // (y, -x)
        case 2: k1 = y; k2 = x; break;                        // (y, x)
        case 5: k1 = x; k2 = -y; break;                       // reverse of (-x, y)
        case 6: k1 = y; k2 = (task == 2) ? -x : 0.0f; break;  // straight: y; right: reverse of (-y, x)
        case 8: k1 = y; k2 = x; break;                        // reverse of (-y, -x)
        case 10: k1 = -x; k2 = 0.0f; break;                   // -x
        default: k1 = y; k2 = 0.0f; break;                    // du, ud: y
    }
}
__device__ __forceinline__ float4 veh_fill_value(int cls) {
    switch (cls) {                                            // mode2fillvalue, E2E:440-447
        case 0: return make_float4(CE2E_LW / 2.0f, -55.0f, 0.0f, 90.0f);
        case 1: return make_float4(CE2E_LW * 1.5f, -55.0f, 0.0f, 90.0f);
        case 2: return make_float4(CE2E_LW * 2.5f, -55.0f, 0.0f, 90.0f);
        case 5: return make_float4(40.0f, CE2E_LW * 2.5f, 0.0f, 180.0f);
        case 6: return make_float4(-CE2E_LW / 2.0f, 45.0f, 0.0f, -90.0f);
        case 7: return make_float4(-CE2E_LW * 1.5f, 45.0f, 0.0f, -90.0f);
        case 8: return make_float4(-CE2E_LW * 2.5f, 45.0f, 0.0f, -90.0f);
        default: return make_float4(-45.0f, -CE2E_LW * 1.5f, 0.0f, 0.0f);       // lr
    }
}

__global__ void k_select_vehicles(const __grid_constant__ SelectSpec spec, int task,
                                  const float *__restrict__ veh_all, const int8_t *__restrict__ route_class,
                                  int N, const float *__restrict__ ego_xy, int v_light,
                                  const int8_t *__restrict__ virtual_red, float *__restrict__ out, int64_t ld,
                                  int64_t B) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * spec.n_modes) return;
    const int64_t i = t / spec.n_modes;
    const int m = (int)(t - i * spec.n_modes);
    const int cls = spec.cls[m], num = spec.num[m];
    const float ex = ego_xy[2 * i], ey = ego_xy[2 * i + 1];
    // best two candidates by (k1, k2, list index)
    float b1[2] = {CUDART_INF_F, CUDART_INF_F}, b2[2] = {CUDART_INF_F, CUDART_INF_F};
    float4 bv[2];
    int nb = 0;
    bv[0] = bv[1] = veh_fill_value(cls);
    auto consider = [&](float4 v) {
        if (!veh_in_range(cls, task, v.x, v.y, ex, ey)) return;
        float k1, k2;
        veh_sort_key(cls, task, v.x, v.y, k1, k2);
        // strict "<": an equal key never displaces an earlier list entry (stable sort)
        const bool lt0 = nb < 1 || k1 < b1[0] || (k1 == b1[0] && k2 < b2[0]);
        const bool lt1 = nb < 2 || k1 < b1[1] || (k1 == b1[1] && k2 < b2[1]);
        if (lt0) {
            b1[1] = b1[0]; b2[1] = b2[0]; bv[1] = bv[0];
            b1[0] = k1; b2[0] = k2; bv[0] = v;
        } else if (lt1) {
            b1[1] = k1; b2[1] = k2; bv[1] = v;
        }
        nb = min(nb + 1, 2);
    };
    const float *va = veh_all + i * (int64_t)N * 4;
    const int8_t *rc = route_class + i * (int64_t)N;
    for (int j = 0; j < N; ++j)
        if (rc[j] == cls) consider(make_float4(va[4 * j], va[4 * j + 1], va[4 * j + 2], va[4 * j + 3]));
    // virtual stopped vehicles at the stop line under a red light (E2E:386-390), appended last
    if (task != 2 && (cls == 0 || cls == 1) && ey < -CE2E_HALF && (v_light != 0 || (virtual_red && virtual_red[i])))
        consider(make_float4(cls == 0 ? CE2E_LW / 2.0f : CE2E_LW * 1.5f, -CE2E_HALF + 2.5f, 0.0f, 90.0f));
    const float4 fill = veh_fill_value(cls);
    float *o = out + i * ld + 4 * spec.slot0[m];
    for (int k = 0; k < num; ++k) {
        const float4 v = (k < nb) ? bv[k] : fill;
        o[4 * k] = v.x; o[4 * k + 1] = v.y; o[4 * k + 2] = v.z; o[4 * k + 3] = v.w;
    }
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

int launch_env_done(int task, const float *obs, int64_t ld, const float *act_scaled, int V, int veh_off,
                    int v_light, int8_t *done, int64_t B, cudaStream_t st) {
    if (V > 0 && aligned16(obs + veh_off) && ld % 4 == 0) {
        const int64_t n_tiles = (B + RPW - 1) / RPW;
        k_env_done<true><<<blocks_for(n_tiles, DONE_WARPS), DONE_WARPS * 32, 0, st>>>(
            make_dyn_consts(1.0 / 10.0), task, obs, ld, act_scaled, V, veh_off, v_light, done, B, DoneReset());
    } else {
        k_env_done<false><<<blocks_for(B, 128), 128, 0, st>>>(make_dyn_consts(1.0 / 10.0), task, obs, ld, act_scaled,
                                                             V, veh_off, v_light, done, B, DoneReset());
    }
    return after_launch("k_env_done");
}

int model_step_common(const ce2e_paths *paths, int task, int path_index, const int32_t *ref_idx,
                      const float *obs_in, int64_t ld_in, const float *act,
                      const ce2e_turn_classes *turn, int V_in, int V_out, int n_future,
                      float *obs_out, int64_t ld_out, float *out5, float *dict16,
                      float *act_scaled_out, int64_t B, int flags, void *stream, int horizon = 0) {
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
        int cur_dev = -1;
        CE2E_CUDA(cudaGetDevice(&cur_dev));
        if (cur_dev != paths->device)
            return fail(CE2E_ERR_PATH, "path tables live on device %d, current device is %d", paths->device, cur_dev);
        P.pv = make_view(paths);
        P.gv = make_grid_view(paths);
        for (int j = 0; j < V_out; ++j) {
            const int tc = turn ? turn->tc[j] : 0;
            P.turn_rs[j] = tc > 0 ? CE2E_R_LEFT : (tc < 0 ? -CE2E_R_RIGHT : 1.0f);
            P.turn_rr[j] = tc > 0 ? host_consts().inv_r_left : (tc < 0 ? -host_consts().inv_r_right : 0.0f);
            P.turn_half[j] = tc != 0 ? CE2E_HALF : -1.0f;
        }
    } else {
        P.pv.n_paths = 0;
        P.pv.stride = 0;
    }
    P.dyn = make_dyn_consts(1.0 / 10.0);          // prediction(..., base_frequency = 10.), DM:387
    P.obs_in = obs_in; P.obs_out = obs_out; P.act = act; P.ref_idx = ref_idx; P.out5 = out5;
    P.dict16 = dict16; P.act_scaled_out = act_scaled_out;
    P.ld_in = ld_in; P.ld_out = ld_out; P.B = B;
    P.task = task; P.path_index = path_index; P.V_in = V_in; P.V_out = V_out; P.n_future = n_future;
    P.horizon = horizon;
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
int ce2e_set_fast_trig(int enable) {
    const int old = g_fast_trig;
    g_fast_trig = enable != 0;
    return old;
}
int ce2e_set_tma(int mode) {
    const int old = g_use_tma;
    g_use_tma = mode < 0 ? 0 : (mode > 3 ? 1 : mode);
    return old;
}
int ce2e_config_set(const ce2e_config *cfg) {
    static const ce2e_config defaults = {4.8, 2.0, 3.75, 3, 50.0, 8.0, 0.05, 0.8, 30.0, 0.02, 5.0, 0.05};
    const ce2e_config c = cfg ? *cfg : defaults;
    if (!(c.L > c.W && c.W > 0 && c.lane_width > 0 && c.lane_number >= 1 && c.crossroad_size > 0))
        return fail(CE2E_ERR_SHAPE, "bad geometry in ce2e_config");
    Ce2eConsts k;
    // the reference's Python doubles, rounded to fp32 where TF would (when they meet a tensor)
    k.lws = (float)((c.L - c.W) / 2);
    k.half = (float)(c.crossroad_size / 2);
    k.lw = (float)c.lane_width;
    k.lw2 = (float)(2 * c.lane_width);
    k.lw3 = (float)(c.lane_width * c.lane_number);
    k.exp_v = (float)c.expected_v;
    k.r_left = (float)(c.crossroad_size / 2 + 0.5 * c.lane_width);
    k.r_right = (float)(c.crossroad_size / 2 - 2.5 * c.lane_width);
    if (!(k.r_right > 0)) return fail(CE2E_ERR_SHAPE, "crossroad_size / 2 - 2.5 lane_width must be positive");
    k.inv_r_left = 1.0f / k.r_left;
    k.inv_r_right = 1.0f / k.r_right;
    k.w_v = (float)c.w_devi_v; k.w_y = (float)c.w_devi_y; k.w_phi = (float)c.w_devi_phi;
    k.w_yaw = (float)c.w_punish_yaw_rate; k.w_steer = (float)c.w_punish_steer; k.w_ax = (float)c.w_punish_a_x;
    int n_dev = 0, cur = 0;
    if (cudaGetDeviceCount(&n_dev) == cudaSuccess && n_dev > 0) {      // every device's constant bank
        CE2E_CUDA(cudaGetDevice(&cur));
        for (int d = 0; d < n_dev; ++d) {
            CE2E_CUDA(cudaSetDevice(d));
            CE2E_CUDA(cudaMemcpyToSymbol(c_consts, &k, sizeof(k)));
        }
        CE2E_CUDA(cudaSetDevice(cur));
    } else {
        cudaGetLastError();
    }
    host_consts() = k;
    g_config = c;
    return CE2E_OK;
}
int ce2e_config_get(ce2e_config *out) {
    if (!out) return fail(CE2E_ERR_NULL, "NULL argument");
    *out = g_config;
    return CE2E_OK;
}
int ce2e_last_step_kernel(void) { return g_last_step_kernel; }
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
    std::vector<float> ph((((size_t)n_paths * stride + 3) & ~(size_t)3), 0.f);   // padded to 16 B (bulk copy)
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
    // candidate grids of find_closest_point (ce2e_grid.h), one per path
    std::vector<uint32_t> all_cells;
    for (int i = 0; i < n_paths; ++i) {
        const int n10 = h->N10[i];
        std::vector<float> wx(n10), wy(n10);
        for (int k = 0; k < n10; ++k) { wx[k] = xs[i][10 * k]; wy[k] = ys[i][10 * k]; }
        std::vector<uint32_t> cells;
        h->cell_off[i] = (int)all_cells.size();
        if (build_candidate_grid(wx.data(), wy.data(), n10, h->grid[i], cells))
            all_cells.insert(all_cells.end(), cells.begin(), cells.end());
        else
            h->grid[i].nx = h->grid[i].ny = 0;
    }
    if (all_cells.empty()) all_cells.push_back(0u);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->cells, sizeof(uint32_t) * all_cells.size());
    if (e == cudaSuccess) e = cudaMemcpy(h->cells, all_cells.data(), sizeof(uint32_t) * all_cells.size(), cudaMemcpyHostToDevice);
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
    if (h->cells) cudaFree(h->cells);
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
                            const float *ys, int ratio, int brute_force, int64_t *idx_out,
                            float *pts_out, int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!paths || !xs || !ys) return fail(CE2E_ERR_NULL, "NULL argument");
    if (path_index < 0 || path_index >= paths->n_paths)
        return fail(CE2E_ERR_PATH, "path_index %d outside [0, %d)", path_index, paths->n_paths);
    if (ratio < 1) return fail(CE2E_ERR_SHAPE, "ratio %d < 1", ratio);
    if (ratio == 10 && !brute_force) {
        k_closest10<<<blocks_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
            make_view(paths), make_grid_view(paths), path_index, xs, ys, idx_out, pts_out, B);
        return after_launch("k_closest10");
    }
    k_closest<<<blocks_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
        paths->full[path_index], paths->L[path_index], ratio, xs, ys, idx_out, pts_out, B);
    return after_launch("k_closest");
}

int ce2e_grid_build_host(const float *wx, const float *wy, int32_t n, float *spec5, uint32_t *cells,
                         int64_t cells_cap) {
    if (!wx || !wy || !spec5) return fail(CE2E_ERR_NULL, "NULL argument");
    GridSpec g;
    std::vector<uint32_t> c;
    if (!build_candidate_grid(wx, wy, n, g, c)) return fail(CE2E_ERR_SHAPE, "no grid for this table");
    spec5[0] = g.x0; spec5[1] = g.y0; spec5[2] = g.inv_h; spec5[3] = (float)g.nx; spec5[4] = (float)g.ny;
    if (cells) {
        if ((int64_t)c.size() > cells_cap) return fail(CE2E_ERR_SHAPE, "cells buffer too small");
        memcpy(cells, c.data(), c.size() * sizeof(uint32_t));
    }
    return CE2E_OK;
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
        make_view(paths), make_grid_view(paths), paths->task, path_index, ref_idx, xs, ys, phis, vs, n_future,
        out, ld_out, B);
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

int ce2e_env_step(const ce2e_paths *paths, const int32_t *ref_idx, const float *obs_in, int64_t ld_in,
                  const float *act_norm, const ce2e_turn_classes *turn, int V, int n_future, int v_light,
                  float *obs_out, int64_t ld_out, float *out5, float *dict16, float *act_scaled_out,
                  int8_t *done_out, int64_t B, void *stream) {
    if (!paths) return fail(CE2E_ERR_NULL, "paths handle is NULL");
    if (B > 0 && (!ref_idx || !act_scaled_out || !done_out)) return fail(CE2E_ERR_NULL, "NULL argument");
    int rc = model_step_common(paths, paths->task, 0, ref_idx, obs_in, ld_in, act_norm, turn, V, V, n_future,
                               obs_out, ld_out, out5, dict16, act_scaled_out, B,
                               F_REWARD | F_NEXT | F_ACT_NORM | F_GYM_EGO, stream);
    if (rc || B == 0) return rc;
    return launch_env_done(paths->task, obs_out, ld_out, act_scaled_out, V, 6 + 3 * (n_future + 1), v_light, done_out,
                           B, (cudaStream_t)stream);
}

static int make_reset_params(const ce2e_paths *paths, uint64_t seed, int fixed_path, int V, int n_future, int64_t ld,
                             ResetParams &R) {
    if (V < 0 || V > CE2E_MAX_VEH || n_future < 0 || ld < 6 + 3 * (n_future + 1) + 4 * V)
        return fail(CE2E_ERR_SHAPE, "bad V / n_future / ld");
    if (fixed_path >= paths->n_paths) return fail(CE2E_ERR_PATH, "fixed_path %d outside [0, %d)", fixed_path, paths->n_paths);
    memset(&R, 0, sizeof(R));
    R.pv = make_view(paths);
    R.gv = make_grid_view(paths);
    for (int i = 0; i < paths->n_paths; ++i) R.full[i] = paths->full[i];
    R.task = paths->task; R.fixed_path = fixed_path < 0 ? -1 : fixed_path; R.V = V; R.n_future = n_future;
    R.span = paths->task == 0 ? 900 + 500 : (paths->task == 1 ? 1200 + 500 : 420 + 500);    // E2E:473-478
    R.seed_lo = (uint32_t)seed; R.seed_hi = (uint32_t)(seed >> 32);
    return CE2E_OK;
}

int ce2e_env_reset(const ce2e_paths *paths, uint64_t seed, int32_t *episode, const int8_t *done, int fixed_path,
                   float *obs, int64_t ld, int32_t *ref_idx, int8_t *virtual_red, int V, int n_future, int64_t B,
                   void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!paths || !episode || !obs || !ref_idx) return fail(CE2E_ERR_NULL, "NULL argument");
    ResetParams R;
    if ((rc = make_reset_params(paths, seed, fixed_path, V, n_future, ld, R))) return rc;
    k_env_reset<<<blocks_for((B + 31) / 32, 8), 256, 0, (cudaStream_t)stream>>>(R, episode, done, obs, ld, ref_idx, virtual_red,
                                                                                  nullptr, B);
    return after_launch("k_env_reset");
}

int ce2e_env_step_reset(const ce2e_paths *paths, int32_t *ref_idx, const float *obs_in, int64_t ld_in,
                        const float *act_norm, const ce2e_turn_classes *turn, int V, int n_future, int v_light,
                        float *obs_out, int64_t ld_out, float *out5, float *dict16, float *act_scaled_out,
                        int8_t *done_out, uint8_t *done_flag_out, uint64_t seed, int32_t *episode, int fixed_path,
                        int8_t *virtual_red, int64_t B, void *stream) {
    if (!paths) return fail(CE2E_ERR_NULL, "paths handle is NULL");
    if (B > 0 && (!ref_idx || !act_scaled_out || !done_out || !episode)) return fail(CE2E_ERR_NULL, "NULL argument");
    int rc = model_step_common(paths, paths->task, 0, ref_idx, obs_in, ld_in, act_norm, turn, V, V, n_future,
                               obs_out, ld_out, out5, dict16, act_scaled_out, B,
                               F_REWARD | F_NEXT | F_ACT_NORM | F_GYM_EGO, stream);
    if (rc || B == 0) return rc;
    const int veh_off = 6 + 3 * (n_future + 1);
    DoneReset X;
    memset(&X, 0, sizeof(X));
    if ((rc = make_reset_params(paths, seed, fixed_path, V, n_future, ld_out, X.R))) return rc;
    X.episode = episode; X.ref_idx = ref_idx; X.virtual_red = virtual_red; X.done_flag = done_flag_out; X.obs_w = obs_out;
    cudaStream_t st = (cudaStream_t)stream;
    if (V > 0 && aligned16(obs_out + veh_off) && ld_out % 4 == 0) {
        // done logic and the restart of the finished rows in ONE launch
        const int64_t n_tiles = (B + RPW - 1) / RPW;
        k_env_done<true, true><<<blocks_for(n_tiles, DONE_WARPS), DONE_WARPS * 32, 0, st>>>(
            make_dyn_consts(1.0 / 10.0), paths->task, obs_out, ld_out, act_scaled_out, V, veh_off, v_light, done_out, B, X);
        return after_launch("k_env_done");
    }
    if ((rc = launch_env_done(paths->task, obs_out, ld_out, act_scaled_out, V, veh_off, v_light, done_out, B, st))) return rc;
    k_env_reset<<<blocks_for((B + 31) / 32, 8), 256, 0, st>>>(X.R, episode, done_out, obs_out, ld_out, ref_idx, virtual_red,
                                                              done_flag_out, B);
    return after_launch("k_env_reset");
}

void ce2e_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    const Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}

int ce2e_judge_done(int task, const float *obs, int64_t ld, const float *act_scaled, int V, int n_future,
                    int v_light, int8_t *done_out, int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if ((rc = check_task(task))) return rc;
    if (B == 0) return CE2E_OK;
    if (!obs || !act_scaled || !done_out) return fail(CE2E_ERR_NULL, "NULL argument");
    if (V < 0 || V > CE2E_MAX_VEH || n_future < 0 || ld < 6 + 3 * (n_future + 1) + 4 * V)
        return fail(CE2E_ERR_SHAPE, "bad V / n_future / ld");
    return launch_env_done(task, obs, ld, act_scaled, V, 6 + 3 * (n_future + 1), v_light, done_out, B,
                           (cudaStream_t)stream);
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

int ce2e_rollout_step_backward(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                               const float *obs_in, int64_t ld_in, const float *act_norm, int V_in,
                               int n_future, const float *g_next, int64_t ld_gnext, const float *g_out5,
                               float *g_obs, int64_t ld_gobs, float *g_act, int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if (B == 0) return CE2E_OK;
    if (!paths || !obs_in || !act_norm || !g_next || !g_out5 || !g_obs || !g_act)
        return fail(CE2E_ERR_NULL, "NULL argument");
    const int n_cols = 6 + 3 * (n_future + 1);
    if (V_in < 0 || V_in > CE2E_MAX_VEH || n_future < 0 || ld_in < n_cols + 4 * V_in || ld_gnext < n_cols ||
        ld_gobs < n_cols)
        return fail(CE2E_ERR_SHAPE, "bad V / n_future / ld");
    if (!ref_idx && (path_index < 0 || path_index >= paths->n_paths))
        return fail(CE2E_ERR_PATH, "path_index %d outside [0, %d)", path_index, paths->n_paths);
    if (aligned16(obs_in + n_cols) && ld_in % 4 == 0 && V_in > 0) {
        const int64_t n_tiles = (B + RPW - 1) / RPW;
        k_model_step_bwd<true><<<blocks_for(n_tiles, TILED_WARPS), TILED_WARPS * 32, 0, (cudaStream_t)stream>>>(
            make_view(paths), make_grid_view(paths), make_dyn_consts(1.0 / 10.0), paths->task, path_index, ref_idx,
            obs_in, ld_in, act_norm, V_in, n_future, g_next, ld_gnext, g_out5, g_obs, ld_gobs, g_act, B);
    } else {
        k_model_step_bwd<false><<<blocks_for(B, 128), 128, 0, (cudaStream_t)stream>>>(
            make_view(paths), make_grid_view(paths), make_dyn_consts(1.0 / 10.0), paths->task, path_index, ref_idx,
            obs_in, ld_in, act_norm, V_in, n_future, g_next, ld_gnext, g_out5, g_obs, ld_gobs, g_act, B);
    }
    return after_launch("k_model_step_bwd");
}

int ce2e_select_vehicles(int task, const float *veh_all, const int8_t *route_class, int N,
                         const float *ego_xy, int v_light, const int8_t *virtual_red, float *out,
                         int64_t ld_out, int64_t B, void *stream) {
    int rc;
    if ((rc = check_batch(B))) return rc;
    if ((rc = check_task(task))) return rc;
    if (B == 0) return CE2E_OK;
    if (!ego_xy || !out || (N > 0 && (!veh_all || !route_class))) return fail(CE2E_ERR_NULL, "NULL argument");
    SelectSpec spec;
    memset(&spec, 0, sizeof(spec));
    // VEHICLE_MODE_DICT (EU:21-23)
    static const int cls_left[4] = {0, 1, 7, 8}, num_left[4] = {2, 2, 2, 2};
    static const int cls_straight[5] = {0, 1, 7, 5, 6}, num_straight[5] = {1, 2, 2, 2, 2};
    static const int cls_right[3] = {2, 6, 10}, num_right[3] = {1, 2, 2};
    const int *cls = task == 0 ? cls_left : task == 1 ? cls_straight : cls_right;
    const int *num = task == 0 ? num_left : task == 1 ? num_straight : num_right;
    spec.n_modes = task == 0 ? 4 : task == 1 ? 5 : 3;
    int slot = 0;
    for (int m = 0; m < spec.n_modes; ++m) {
        spec.cls[m] = cls[m]; spec.num[m] = num[m]; spec.slot0[m] = slot;
        slot += num[m];
    }
    if (N < 0 || ld_out < 4 * slot) return fail(CE2E_ERR_SHAPE, "bad N / ld_out (need %d columns)", 4 * slot);
    k_select_vehicles<<<blocks_for(B * spec.n_modes, 128), 128, 0, (cudaStream_t)stream>>>(
        spec, task, veh_all, route_class, N, ego_xy, v_light, virtual_red, out, ld_out, B);
    return after_launch("k_select_vehicles");
}

int ce2e_rollout_horizon(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                         const float *obs_in, int64_t ld_in, const float *act_tape,
                         const ce2e_turn_classes *turn, int V, int n_future, int H, float *obs_out,
                         int64_t ld_out, float *out5, int64_t B, void *stream) {
    if (!paths) return fail(CE2E_ERR_NULL, "paths handle is NULL");
    if (H < 1 || H > (1 << 20)) return fail(CE2E_ERR_SHAPE, "bad horizon %d", H);
    if (V > 4 * 2 * FUSED_MAX_CHUNKS)
        return fail(CE2E_ERR_SHAPE, "the horizon-fused kernel keeps at most %d vehicles per row on chip",
                    4 * 2 * FUSED_MAX_CHUNKS);
    return model_step_common(paths, paths->task, path_index, ref_idx, obs_in, ld_in, act_tape, turn, V, V,
                             n_future, obs_out, ld_out, out5, nullptr, nullptr, B,
                             F_REWARD | F_NEXT | F_ACT_NORM, stream, H);
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
