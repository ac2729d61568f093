/* ce2e.h -- C ABI of libce2e.so: the CrossroadEnd2end model hot path on B200 (sm_100a).
 *
 * This is the drop-in boundary for the data-parallel path of idthanm/env_build
 * (SURVEY.md section 8b).  Every entry point names the reference interface it
 * replaces; citations are file:line into the reference checkout
 * (DM = dynamics_and_models.py, E2E = endtoend.py, EU = endtoend_env_utils.py).
 *
 * Conventions
 *   - All tensor pointers are DEVICE pointers to fp32 (or int32/int64 where stated),
 *     owned by the caller.  The library never allocates or frees caller buffers,
 *     never synchronises and enqueues all work on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream), so every call is CUDA-graph capturable.
 *   - Observation rows are AoS fp32 with a leading dimension `ld` (floats between
 *     consecutive rows, ld >= D), BLAS style, so a caller may keep rows padded for
 *     16-byte alignment of the vehicle block.  Row layout (E2E:300, DM:189-194):
 *        [v_x v_y r x y phi_deg | d_y d_phi_deg d_v | (dx dy dphi)*n | (x y v phi_deg)*V]
 *   - Every function returns 0 on success or a negative CE2E_ERR_* code and stores a
 *     thread-local message readable with ce2e_last_error().
 *   - Kernels never trap on NaN/Inf; they propagate like the reference's TF ops.
 *   - fp32 arithmetic follows the reference's expression trees with one rounding per
 *     op and no FMA contraction; sin/cos/atan are CUDA's
 *     <= 2 ulp routines.
 */
#ifndef CE2E_H_
#define CE2E_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CE2E_VERSION 100          /* 0.1.0 */
#define CE2E_MAX_PATHS 4
#define CE2E_MAX_VEH 256          /* vehicles per observation row */

#define CE2E_OK 0
#define CE2E_ERR_NULL (-1)        /* required pointer is NULL */
#define CE2E_ERR_SHAPE (-2)       /* bad B / V / n / ld / length */
#define CE2E_ERR_TASK (-3)        /* task not in {left, straight, right} (DM:274, DM:668, DM:747 asserts) */
#define CE2E_ERR_PATH (-4)        /* path index outside the handle's path list */
#define CE2E_ERR_CUDA (-5)        /* CUDA runtime error (message has the cudaError string) */
#define CE2E_ERR_NOMEM (-6)

/* training_task of EnvironmentModel / ReferencePath (DM:91, DM:584) */
#define CE2E_TASK_LEFT 0
#define CE2E_TASK_STRAIGHT 1
#define CE2E_TASK_RIGHT 2

/* turn class of one surrounding vehicle = the branch predict_for_a_mode takes (DM:416-421):
 *  +1 for modes dl, rd, ur, lu;  -1 for dr, ru, ul, ld;  0 otherwise.                  */
typedef struct ce2e_turn_classes {
    int8_t tc[CE2E_MAX_VEH];
} ce2e_turn_classes;

/* Opaque device-resident copy of a task's reference-path tables
 * (ReferencePath.path_list, DM:584-700).                                                  */
typedef struct ce2e_paths ce2e_paths;

int ce2e_version(void);
const char *ce2e_last_error(void);

/* Scenario constants the reference bakes in at import: geometry (EU:14-18: L, W, LANE_WIDTH, LANE_NUMBER,
 * CROSSROAD_SIZE; EXPECTED_V) and the reward weights of compute_rewards (DM:297-298).  Defaults are the
 * reference's values.  ce2e_config_set(NULL) restores them; the setting is process wide, applies to every
 * device and to every later launch (launch nothing concurrently).  Reference-path tables are inputs
 * (ce2e_paths_create), so a caller that changes the geometry also passes matching tables.         */
typedef struct ce2e_config {
    double L, W;                 /* ego / vehicle length and width: circle centres at +-(L - W)/2, DM:209 */
    double lane_width;           /* LANE_WIDTH */
    int lane_number;             /* LANE_NUMBER */
    double crossroad_size;       /* CROSSROAD_SIZE */
    double expected_v;           /* EXPECTED_V (delta_v of the tracking error, DM:760) */
    double w_devi_v, w_devi_y, w_devi_phi, w_punish_yaw_rate, w_punish_steer, w_punish_a_x;   /* DM:297-298 */
} ce2e_config;
int ce2e_config_set(const ce2e_config *cfg);
int ce2e_config_get(ce2e_config *out);

/* Process-wide option, default 0.  enable != 0: ce2e_rollout_step takes sin / cos of the SURROUNDING
 * VEHICLES' headings (DM:220-224, DM:409-410) from the special-function unit (abs. error <= 2^-21.4
 * instead of <= 1.5 ulp).  Results stay inside the 1e-5 parity tolerance; everything else is
 * unchanged.  Returns the previous setting.                                                   */
int ce2e_set_fast_trig(int enable);

/* Process-wide option, default 1 (environment CE2E_NO_TMA=1 starts with 0).  mode != 0: the fused
 * step (ce2e_rollout_step / ce2e_env_step; EnvironmentModel.rollout_out, DM:118-126) runs the
 * warp-pair kernel that streams the vehicle block and the ego columns with TMA tensor-map copies
 * whenever the vehicle block of obs_in and obs_out is 16-byte aligned with ld % 4 == 0, V_in == V_out
 * >= 1 and B >= 32; otherwise, and with mode == 0, the cp.async kernel runs.  mode 1 picks the pair's
 * work split from the batch size, 2 / 3 force the overlapped / the balanced split (tests, A/B runs).
 * All variants give bit-identical results.  Returns the previous mode.                          */
int ce2e_set_tma(int mode);

/* Which kernel the calling thread's last fused-step launch (rollout_out, DM:118-126) used: 0 none
 * yet, 1 k_model_step (cp.async staging), 2 k_model_step_pair (TMA tensor-map staging).
 * Diagnostic only.                                                                             */
int ce2e_last_step_kernel(void);

/* Number of CUDA kernels this library has launched from the calling thread since load. */
int64_t ce2e_launch_count(void);

/* ReferencePath.__init__ (DM:584-592): upload n_paths tables of HOST fp32 arrays
 * xs[i], ys[i], phis[i] of length lens[i] (path_list[i], DM:631).  The handle keeps the
 * full tables plus the every-10th-point copy find_closest_point uses (DM:704-706).
 * Synchronous (one-time setup).                                                           */
int ce2e_paths_create(int task, int n_paths, const int32_t *lens, const float *const *xs,
                      const float *const *ys, const float *const *phis, ce2e_paths **out);
int ce2e_paths_destroy(ce2e_paths *paths);

/* EnvironmentModel._action_transformation_for_end2end (DM:128-132; NumPy twin E2E:258-267).
 * act_norm, act_scaled: [B,2] contiguous.                                                 */
int ce2e_action_transform(const float *act_norm, float *act_scaled, int64_t B, void *stream);

/* VehicleDynamics.f_xu(states, actions, tau) (DM:52-83) / .prediction (DM:85-87).
 * states [B,6] (ld_states), actions [B,2] contiguous SCALED actions (steer rad, a_x), tau the
 * Python float the reference passes (rounded to fp32 where TF does),
 * next [B,6] (ld_next); params [B,4] contiguous = (alpha_f, alpha_r, miu_f, miu_r) or NULL.
 * clip_vx != 0 additionally applies ego_predict's v_x = clip(v_x, 0, 35) (DM:386-392).     */
int ce2e_dynamics_step(const float *states, int64_t ld_states, const float *actions, double tau,
                       float *next, int64_t ld_next, float *params, int clip_vx, int64_t B,
                       void *stream);

/* ReferencePath.find_closest_point(xs, ys, ratio) (DM:702-715) on path `path_index`:
 * idx_out [B] int64 = ratio * first-argmin of the squared distance over every ratio-th
 * waypoint; pts_out [3,B] contiguous = (x, y, phi_deg) of that waypoint (indexs2points,
 * DM:726-733).  Either output may be NULL.  With ratio == 10 and brute_force == 0 the scan
 * is restricted to the candidate range of the query's grid cell (same result, see
 * ce2e_grid_build_host); brute_force != 0 evaluates every candidate like the reference.     */
int ce2e_find_closest_point(const ce2e_paths *paths, int path_index, const float *xs,
                            const float *ys, int ratio, int brute_force, int64_t *idx_out,
                            float *pts_out, int64_t B, void *stream);

/* HOST function (no CUDA): the candidate grid libce2e builds for the n every-10th waypoints
 * (wx, wy) of one path, as used for find_closest_point (DM:702-715).  spec5 = {x0, y0,
 * cells per metre, nx, ny}; cells[iy*nx + ix] = lo | hi << 16 means: for every query point
 * in cell ix = (int)((x - x0) * spec5[2]), iy likewise, the brute-force first-argmin lies in
 * [lo, hi].  cells may be NULL to query the size.                                           */
int ce2e_grid_build_host(const float *wx, const float *wy, int32_t n, float *spec5, uint32_t *cells,
                         int64_t cells_cap);

/* ReferencePath.indexs2points (DM:726-733) / future_n_data (DM:717-724).
 * n_future == 0: pts_out [3,B] = points at clamp(idx).  n_future > 0: pts_out
 * [n_future,3,B] = the preview points idx+80k clamped to L-2.                              */
int ce2e_index_points(const ce2e_paths *paths, int path_index, const int64_t *idx, int n_future,
                      float *pts_out, int64_t B, void *stream);

/* ReferencePath.tracking_error_vector(xs, ys, phis, vs, n) (DM:735-770).
 * ref_idx == NULL: every row uses path `path_index` (mode != 'training', DM:334-339).
 * ref_idx != NULL ([B] int32): row i uses path ref_idx[i]; a value outside the path
 * list yields zeros (mode == 'training', DM:340-353).  out: [B, 3(n+1)] with ld_out.       */
int ce2e_tracking_error(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                        const float *xs, const float *ys, const float *phis, const float *vs,
                        int n_future, float *out, int64_t ld_out, int64_t B, void *stream);

/* EnvironmentModel.compute_rewards(obses, actions) (DM:186-320).  actions = SCALED [B,2].
 * out5 [5,B] contiguous = rewards, punish_term_for_training, real_punish_term,
 * veh2veh4real, veh2road4real.  dict16 [16,B] or NULL = reward_dict in the key order of
 * DM:302-318 (punish_steer, punish_a_x, punish_yaw_rate, devi_v, devi_y, devi_phi,
 * scaled_punish_steer, scaled_punish_a_x, scaled_punish_yaw_rate, scaled_devi_v,
 * scaled_devi_y, scaled_devi_phi, veh2veh4training, veh2road4training, veh2veh4real,
 * veh2road4real).                                                                         */
int ce2e_compute_rewards(int task, const float *obs, int64_t ld, const float *actions, int V,
                         int n_future, float *out5, float *dict16, int64_t B, void *stream);

/* EnvironmentModel.compute_next_obses(obses, actions) (DM:322-358): the next-observation half
 * of ce2e_rollout_step on its own; actions = SCALED [B,2].                                 */
int ce2e_compute_next_obses(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                            const float *obs_in, int64_t ld_in, const float *actions,
                            const ce2e_turn_classes *turn, int V_in, int V_out, int n_future,
                            float *obs_out, int64_t ld_out, int64_t B, void *stream);

/* One step of a batch of SUMO-free CrossroadEnd2end environments (E2E:132-144) whose surrounding
 * traffic follows the analytic model: action scaling (E2E:258-267) -> compute_reward on obs_in
 * (E2E:501-507 = DM:186-320; out5 / dict16 as in ce2e_compute_rewards) -> _get_next_ego_state
 * (E2E:269-283: f_xu at 10 Hz, v_x floored at 0, heading wrapped to (-180, 180]) -> vehicles
 * advanced by veh_predict (DM:394-427, standing in for Traffic.sim_step) -> _get_obs tracking
 * part on the row's path ref_idx[i] (E2E:285-303) -> _judge_done on the new state (E2E:200-256,
 * collision test of traffic.py:263-295).  done_out [B] int8: 0 not_done_yet, 1 collision,
 * 2 break_road_constrain, 3 deviate_too_much, 4 break_stability, 5 break_red_light, 6 good_done.
 * act_scaled_out [B,2] is required (the stability bound needs the scaled a_x).               */
int ce2e_env_step(const ce2e_paths *paths, const int32_t *ref_idx, const float *obs_in, int64_t ld_in,
                  const float *act_norm, const ce2e_turn_classes *turn, int V, int n_future, int v_light,
                  float *obs_out, int64_t ld_out, float *out5, float *dict16, float *act_scaled_out,
                  int8_t *done_out, int64_t B, void *stream);

/* CrossroadEnd2end.reset / _reset_init_state (E2E:99-127, E2E:472-499) for the batched environment, on
 * the device: every row i with done[i] != 0 (done == NULL: every row) starts a new episode -- waypoint
 * index int(u * span) + 700 on its path (E2E:473-478), ego = (8 u', 0, 0, x, y, phi) (E2E:480-499),
 * tracking columns re-projected (E2E:293-297), the V vehicle slots refilled from the synthetic traffic
 * distribution (the reference asks SUMO, traffic.py:151-195 -- out of scope), ref_idx[i] = fixed_path, or
 * a uniformly drawn path if fixed_path < 0.  Draws are Philox4x32-10 at counter (i, episode[i], block)
 * under key `seed`; episode[i] ([B] int32, caller-initialised) is incremented for the rows reset.
 * virtual_red ([B] int8 or NULL) receives the 10 % virtual-red-light flag of E2E:120-124.
 * No host synchronisation: capturable together with ce2e_env_step.                               */
int ce2e_env_reset(const ce2e_paths *paths, uint64_t seed, int32_t *episode, const int8_t *done,
                   int fixed_path, float *obs, int64_t ld, int32_t *ref_idx, int8_t *virtual_red, int V,
                   int n_future, int64_t B, void *stream);

/* CrossroadEnd2end.step (E2E:132-144) with the reference's `done -> reset()` loop folded in (E2E:99-127):
 * ce2e_env_step followed by ce2e_env_reset of the rows whose done code is non-zero, as two launches (fused
 * model step; done logic + restart).  ref_idx is read by the step and rewritten for the restarted rows;
 * done_out keeps the codes of the step just taken, done_flag_out ([B] bytes 0 / 1, or NULL) = code != 0;
 * obs_out holds the next observation, or the first observation of the next episode for restarted rows.
 * Arguments as in ce2e_env_step and ce2e_env_reset.  No host synchronisation: capturable.          */
int ce2e_env_step_reset(const ce2e_paths *paths, int32_t *ref_idx, const float *obs_in, int64_t ld_in,
                        const float *act_norm, const ce2e_turn_classes *turn, int V, int n_future, int v_light,
                        float *obs_out, int64_t ld_out, float *out5, float *dict16, float *act_scaled_out,
                        int8_t *done_out, uint8_t *done_flag_out, uint64_t seed, int32_t *episode, int fixed_path,
                        int8_t *virtual_red, int64_t B, void *stream);

/* HOST function: one Philox4x32-10 block, the generator behind ce2e_env_reset's draws (E2E:473, E2E:482
 * use np.random.random()), exported so that tests can pin it against known-answer vectors.        */
void ce2e_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* CrossroadEnd2end._judge_done (E2E:200-256) with Traffic.collision_check (traffic.py:263-295) on
 * observations obs [B,D] taken AFTER a step; act_scaled [B,2] = the scaled action of that step
 * (its a_x fixes miu_r in the yaw-rate bound, E2E:167).  done_out as in ce2e_env_step.       */
int ce2e_judge_done(int task, const float *obs, int64_t ld, const float *act_scaled, int V, int n_future,
                    int v_light, int8_t *done_out, int64_t B, void *stream);

/* EnvironmentModel.veh_predict (DM:394-427): veh_in/veh_out point at the first vehicle
 * column of each row ([B,4V] with ld_in / ld_out).                                        */
int ce2e_veh_predict(const float *veh_in, int64_t ld_in, const ce2e_turn_classes *turn, int V,
                     float *veh_out, int64_t ld_out, int64_t B, void *stream);

/* EnvironmentModel.rollout_out(actions) (DM:118-126), one fused launch:
 *   action scaling (DM:128-132) -> compute_rewards on obs_in (DM:186-320) ->
 *   compute_next_obses (DM:322-358) = ego_predict + tracking_error_vector on the row's
 *   path + veh_predict.
 * obs_in [B, 6+3(n+1)+4*V_in] (ld_in); obs_out [B, 6+3(n+1)+4*V_out] (ld_out), V_out <= V_in
 * (the reference predicts only len(VEHICLE_MODE_LIST[task]) vehicles, DM:398-402; pass
 * V_out = V_in with a length-V_in turn list for the general case).  obs_out must not
 * overlap obs_in.  act_norm [B,2] in [-1,1]; ref_idx/path_index as in ce2e_tracking_error;
 * out5 [5,B] as in ce2e_compute_rewards; act_scaled_out [B,2] or NULL receives
 * self.actions (DM:120).                                                                  */
int ce2e_rollout_step(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                      const float *obs_in, int64_t ld_in, const float *act_norm,
                      const ce2e_turn_classes *turn, int V_in, int V_out, int n_future,
                      float *obs_out, int64_t ld_out, float *out5, float *act_scaled_out,
                      int64_t B, void *stream);

/* Vector-Jacobian product of EnvironmentModel.rollout_out (DM:118-126), i.e. what TensorFlow's
 * autodiff gives the reference's model-based trainer: vehicle columns are constants
 * (tf.stop_gradient, DM:195 / 331 / 402), the closest-waypoint index is an integer (tf.argmin,
 * DM:714), tf.where / tf.clip_by_value pass the selected branch.  obs_in / act_norm / ref_idx /
 * path_index: the inputs of the forward ce2e_rollout_step.  g_next [B, 6+3(n+1)] (ld_gnext): upstream
 * gradient w.r.t. the next ego + tracking columns; g_out5 [5,B]: w.r.t. rewards, punish_term_for_
 * training, real_punish_term, veh2veh4real, veh2road4real.  Outputs: g_obs [B, 6+3(n+1)] (ld_gobs):
 * gradient w.r.t. the ego + tracking columns of obs_in; g_act [B,2]: w.r.t. the normalised actions. */
int ce2e_rollout_step_backward(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                               const float *obs_in, int64_t ld_in, const float *act_norm, int V_in,
                               int n_future, const float *g_next, int64_t ld_gnext, const float *g_out5,
                               float *g_obs, int64_t ld_gobs, float *g_act, int64_t B, void *stream);

/* CrossroadEnd2end._construct_veh_vector_short (E2E:340-464) for B scenes: veh_all [B,N,4] =
 * (x, y, v, phi_deg) of the vehicles around each ego, route_class [B,N] int8 = 0 dl, 1 du, 2 dr,
 * 3 rd, 4 rl, 5 ru, 6 ur, 7 ud, 8 ul, 9 lu, 10 lr, 11 ld (the lists of E2E:354; other values are
 * ignored), ego_xy [B,2], v_light as in the reference (0 = green), virtual_red [B] int8 or NULL =
 * self.virtual_red_light_vehicle (E2E:120-124).  out [B, 4*VEH_NUM[task]] (ld_out): per route
 * class of VEHICLE_MODE_DICT[task] (EU:21-23) the vehicles passing the range filter (E2E:393-411),
 * stably sorted by the class's key (E2E:414-428), cut / padded to the class's count with the fill
 * vehicles of E2E:440-447.                                                                    */
int ce2e_select_vehicles(int task, const float *veh_all, const int8_t *route_class, int N,
                         const float *ego_xy, int v_light, const int8_t *virtual_red, float *out,
                         int64_t ld_out, int64_t B, void *stream);

/* H consecutive EnvironmentModel.rollout_out steps (DM:118-126) of an OPEN-LOOP action tape in one
 * launch (the shield / MPC rollouts of hier_decision.py:89-97, multi_ego.py:187-197 with the actions
 * known in advance): act_tape [H,B,2] normalised actions, out5 [H,5,B] the five outputs of every
 * step, obs_out [B,D] the observations after the last step.  The tile's state stays on chip between
 * steps; results are bit-identical to H calls of ce2e_rollout_step.  V <= 32.                   */
int ce2e_rollout_horizon(const ce2e_paths *paths, int path_index, const int32_t *ref_idx,
                         const float *obs_in, int64_t ld_in, const float *act_tape,
                         const ce2e_turn_classes *turn, int V, int n_future, int H, float *obs_out,
                         int64_t ld_out, float *out5, int64_t B, void *stream);

/* EnvironmentModel.ss(obses, actions, lam) (DM:134-184): discrete barrier penalty.
 * obs/next_obs [B,D] rows with V vehicles each (next_obs from ce2e_rollout_step or
 * compute_next_obses); out [B].                                                           */
int ce2e_ss(const float *obs, int64_t ld, const float *next_obs, int64_t ld_next, int V,
            int n_future, double lam, float *out, int64_t B, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CE2E_H_ */
