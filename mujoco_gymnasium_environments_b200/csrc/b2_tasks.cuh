// b2_tasks.cuh -- per-task clip/ctrl, observation, reward, termination and auto-reset fused on the end of the step.
// Each task restates one reference env's order of operations including its index-aliasing
// quirks; all reads of xpos/contacts use the kinematics of the *last forward pass* (SURVEY F9).
#pragma once
#include "b2_engine.cuh"

namespace b2 {

enum { TASK_NONE = 0, TASK_QUADRUPED_PARKOUR = 1, TASK_HUMANOID_DANCING = 2, TASK_HUMANOID_SOCCER = 3, TASK_BIPEDAL_RESCUE = 4, TASK_HUMANOID_CONSTRUCTION = 5, TASK_HUMANOID_MARTIAL_ARTS = 6, TASK_ROBOTIC_ARM_ASSEMBLY = 7 };

// counter-based RNG (splitmix64 finaliser over (seed, env, episode, draw)); documented stream layout in DESIGN.md
__device__ __forceinline__ float rng_uniform(unsigned long long seed, unsigned env, unsigned episode, unsigned draw) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (((unsigned long long)env << 32) ^ ((unsigned long long)episode << 8) ^ draw ^ 0x5851F42D4C957F2Dull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

struct TaskParams {
  int task;
  int ids[16];      // body / joint / actuator ids resolved by the host from names (mj_name2id stand-in)
  float act_lo[40], act_hi[40];
  int aux_i[64]; float aux_f[32];   // task-specific tables (see B2TaskDesc)
};

// ------------------------------------------------------------------------------------------------ quadruped
// quadruped_parkour_env/parkour_env.py: step :356-394, _get_observation :396-468, _calculate_reward :646-725,
// _is_terminated :727-755, reset :314-354, _randomize_obstacles :757-774, _update_dynamic_obstacles :776-795.
// ti: [0] step_count [1] checkpoint bitset (6 checkpoints + 12 obstacles) [2] fall_count [3] stuck_counter [4] episode id
// tf: [0] episode_reward [1] last_position.x [2] max_forward_progress
// ids: [0] torso body, [1..4] foot bodies, [5] platform_slide joint id, [6] pendulum_swing joint id,
//      [7] platform_motor actuator id, [8] pendulum_motor actuator id
struct QuadrupedTask {
  static constexpr int OBS = 95, ACT = 16, FRAME_SKIP = 10, SETTLE = 10, MAX_STEPS = 6000, NTI = 8, NTF = 4, NINJ = 4, KEEP_FRAMES = 0, XFRC_SLOT = -1, COOP_MIN = 1 << 20, ARENA_ROWS = 72, CON_CAP = 48, ARENA_SPAN = 0, MAX_EPB = 4, EPISODE_SLOT = 4, LOCKSTEP = 0, NEWTON_TEAM_ND = 16, ARENA_FLOATS = 0;
  static constexpr int SOLVER = 0;      // compile-time copy of the model's <option solver>; b2_batch_create checks it
  static constexpr bool CONDIM6 = false, RESET_FORWARD = false, PGS_HOIST = true, COLD_PAIRS = false, DYN_ISLANDS = false, CONVEX_PAIRS = false;    // robot and obstacles only ever touch the planes

  template <class EN> __device__ static void apply_action(EN& E, const TaskParams& tp, const float* act, float* act_clipped) {
    for (int i = E.lane; i < ACT; i += 32) {
      float a = clampf(act[i], tp.act_lo[i], tp.act_hi[i]);
      act_clipped[i] = a; E.p_ctrl()[i] = a;
    }
    E.sync();
  }

  template <class EN> __device__ static void reset_state(EN& E, const TaskParams& tp, const BatchView& B, int env, int* ti, float* tf,
                                     const float* inject) {
    E.reset_data();
    if (E.lane == 0) {
      float* q = E.p_qpos();
      q[0] = 2.0f; q[1] = 0.0f; q[2] = 0.6f; q[3] = 1.f; q[4] = 0.f; q[5] = 0.f; q[6] = 0.f;
      unsigned ep = (unsigned)ti[4];
      float u0 = inject ? inject[0] : -1.5f + 3.0f * rng_uniform(B.seed, (unsigned)(B.env_offset + env), ep, 0);
      float u1 = inject ? inject[1] : -1.0f + 2.0f * rng_uniform(B.seed, (unsigned)(B.env_offset + env), ep, 1);
      q[tp.ids[5]] = u0;   // joint id used as a qpos address: lands on bl_knee / bl_ankle
      q[tp.ids[6]] = u1;
      ti[0] = 0; ti[1] = 0; ti[2] = 0; ti[3] = 0; ti[4] = (int)(ep + 1);
      tf[0] = 0.f; tf[1] = 2.0f; tf[2] = 0.f;
      *E.p_time() = 0.f;
    }
    E.sync();
  }

  __device__ static __forceinline__ void obstacle(int k, float& x, float& typ, float& hgt, float& dif) {
    const float X[12] = {8, 16, 24, 30, 36, 44, 50, 58, 72, 78, 88, 92};
    const float H[12] = {0.225f, 0.2f, 0.5f, 0.6f, 0.6f, 0.3f, 0.08f, 0.4f, 0.25f, 0.3f, 0.0f, 0.2f};
    const float D[12] = {0.3f, 0.6f, 0.8f, 0.4f, 0.7f, 0.9f, 0.5f, 0.6f, 0.4f, 0.8f, 1.0f, 1.0f};
    x = X[k]; typ = (float)(k + 1); hgt = H[k]; dif = D[k];
  }

  template <class EN> __device__ static void observe(EN& E, const TaskParams& tp, float* obs) {
    int torso = tp.ids[0];
    for (int i = E.lane; i < OBS; i += 32) {
      float v = 0.f;
      if (i < 16) v = E.p_qpos()[7 + i];
      else if (i < 32) v = E.p_qvel()[6 + i - 16];
      else if (i < 36) v = E.p_qpos()[3 + i - 32];
      else if (i < 42) v = E.p_qvel()[i - 36];
      else if (i < 45) v = E.p_qpos()[i - 42];
      else if (i < 49) {
        // foot *body* id compared with contact *geom* ids
        int fid = tp.ids[1 + i - 45]; int ncon = E.p_misc()[MISC_NCON];
        const int* pc1 = E.PI(DI_pair_cg1); const int* pc2 = E.PI(DI_pair_cg2); const int* gid = E.I(DI_cg_geomid);
        for (int c = 0; c < ncon; c++) {
          int p = __float_as_int(E.x_con()[B2_CON_STRIDE * c + 13]);
          if (gid[pc1[p]] == fid || gid[pc2[p]] == fid) { v = 1.f; break; }
        }
      } else if (i < 61) { int f = (i - 49) / 3, k = (i - 49) % 3; v = E.p_xpos()[3 * tp.ids[1 + f] + k] - E.p_xpos()[3 * torso + k]; }
      else if (i < 85) v = 10.0f;
      else if (i < 93) {
        float x = E.p_xpos()[3 * torso]; int slot = (i - 85) / 4, fld = (i - 85) % 4, found = 0;
        for (int k = 0; k < 12; k++) {
          float ox, ty, hg, df; obstacle(k, ox, ty, hg, df);
          if (ox > x) { if (found == slot) { v = fld == 0 ? ox - x : fld == 1 ? ty : fld == 2 ? hg : df; break; } found++; }
        }
      } else if (i == 94) v = 0.8f;
      obs[i] = v;
    }
  }

  // returns reward; updates ti/tf; sets *terminated
  template <class EN> __device__ static float reward_and_done(EN& E, const TaskParams& tp, const float* act, int* ti, float* tf,
                                          int* terminated, int* truncated) {
    int torso = tp.ids[0];
    float x = E.p_xpos()[3 * torso], y = E.p_xpos()[3 * torso + 1], z = E.p_xpos()[3 * torso + 2];
    float reward = -20.0f;
    float progress = x - tf[1];
    if (progress > 0.f) { reward += progress * 500.0f; tf[2] = fmaxf(tf[2], x); }
    else if (progress < -0.1f) reward -= 100.0f;
    int bits = ti[1];
    for (int k = 0; k < 6; k++) { float cx = 15.0f * (k + 1); if (!(bits >> k & 1) && x >= cx) { bits |= 1 << k; reward += 1000.0f; } }
    for (int k = 0; k < 12; k++) {
      float ox, ty, hg, df; obstacle(k, ox, ty, hg, df);
      if (!(bits >> (6 + k) & 1) && x > ox + 2.0f) { bits |= 1 << (6 + k); reward += 1000.0f + df * 1000.0f; }
    }
    ti[1] = bits;
    if (x >= 98.0f) reward += 5000.0f;
    if (fabsf(E.p_qpos()[3]) > 0.7f) reward += 100.0f;
    int cc = 0;
    {
      int ncon = E.p_misc()[MISC_NCON];
      const int* pc1 = E.PI(DI_pair_cg1); const int* pc2 = E.PI(DI_pair_cg2); const int* gid = E.I(DI_cg_geomid);
      for (int f = 0; f < 4; f++) {
        int fid = tp.ids[1 + f];
        for (int c = 0; c < ncon; c++) {
          int p = __float_as_int(E.x_con()[B2_CON_STRIDE * c + 13]);
          if (gid[pc1[p]] == fid || gid[pc2[p]] == fid) { cc++; break; }
        }
      }
    }
    if (cc >= 1 && cc <= 3) reward += 200.0f;
    float effort = 0.f;
    for (int i = 0; i < ACT; i++) effort += fabsf(act[i]);
    reward -= effort * 0.1f;
    if (z < 0.2f) { reward -= 2000.0f; ti[2] += 1; }
    if (E.p_misc()[MISC_NCON] > 8) reward -= 500.0f;
    if (fabsf(progress) < 0.01f) { ti[3] += 1; if (ti[3] > 100) reward -= 100.0f; }
    else ti[3] = 0;
    tf[1] = x;
    *terminated = (x >= 98.0f) || (z < 0.15f) || (fabsf(y) > 10.0f) || (ti[3] > 1000) || (ti[2] > 3);
    *truncated = ti[0] >= MAX_STEPS;
    ti[0] += 1;
    tf[0] += reward;
    return reward;
  }

  template <class EN> __device__ static void pre_physics(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void after_settle(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void post_physics(EN& E, const TaskParams& tp, int* ti, float*) {
    // _update_dynamic_obstacles: uses the pre-increment step counter; takes effect on the next step
    if (E.lane == 0) {
      float t = (float)ti[0] * 0.01f;
      E.p_ctrl()[tp.ids[7]] = 50.0f * sinf(0.5f * t);
      E.p_ctrl()[tp.ids[8]] = 100.0f * sinf(0.3f * t);
    }
    E.sync();
  }
};

// ------------------------------------------------------------------------------------------------ humanoid dancing
// humanoid_dancing_env/dancing_env.py: step :833-894, reset :763-831, _generate_dance_sequence :896-905,
// _set_initial_pose :907-922, _update_rhythm :924-937, _update_visual_effects :939-955, _check_move_transition :957-973,
// _update_crowd_excitement :975-1002, _update_episode_stats :1004-1026, _get_observation :1028-1120,
// _calculate_reward :1122-1207, _check_termination :1209-1234.
// The scalar clocks the reference keeps in Python floats (time_since_last_beat, combo_multiplier, move_start_time) are
// fp64 here as well: their threshold tests (beat phase 0.1/0.9, int(combo), 0.8 * duration) decide reward branches.
// ti: [0] current_step [1] beat_count [2] current_measure [3] current_move_idx [4] episode id [5] fall_start_step + 1
//     (0 = attribute absent; survives reset) [6] longest_combo [7] len(move_history) [8] last three history
//     entries, 4 bits each [9..11] the 20 moves of dance_sequence, 4 bits each
// tf: [0] performance_score [1] crowd_excitement [2,3] time_since_last_beat (f64) [4,5] combo_multiplier (f64)
//     [6,7] move_start_time (f64) [8..10] spotlight_position (survives reset) [11] energy_used [12] time_on_beat
//     [13..15] xpos[torso] of the last forward pass [16..35] move durations [36..58] prev_joint_vel = qvel[6:]
// ids: [0] torso body [1] right_foot geom [2] left_foot geom [3] dance_floor geom [4] stage geom
struct DancingTask {
  static constexpr int OBS = 94, ACT = 29, FRAME_SKIP = 1, SETTLE = 10, MAX_STEPS = 3600, NTI = 12, NTF = 60, NINJ = 40, KEEP_FRAMES = 1, XFRC_SLOT = -1, COOP_MIN = 1 << 20, ARENA_ROWS = 60, CON_CAP = 32, ARENA_SPAN = 0, MAX_EPB = 5, EPISODE_SLOT = 4, LOCKSTEP = 1, NEWTON_TEAM_ND = 16, ARENA_FLOATS = 0;
  static constexpr int SOLVER = 0;      // compile-time copy of the model's <option solver>; b2_batch_create checks it
  static constexpr bool CONDIM6 = false, RESET_FORWARD = false, PGS_HOIST = false, COLD_PAIRS = false, DYN_ISLANDS = false, CONVEX_PAIRS = true;   // a single kinematic tree
  static constexpr int NJ = 29, NSEQ = 20;
  static constexpr double DT = 0.01667, BEAT = 0.5;

  __device__ static __forceinline__ double& D(float* tf, int k) { return *reinterpret_cast<double*>(tf + k); }
  __device__ static __forceinline__ int seq_move(const int* ti, int k) { return (ti[9 + (k >> 3)] >> (4 * (k & 7))) & 15; }
  __device__ static __forceinline__ float difficulty(int move) {
    const float dif[10] = {1.f, 2.f, 2.f, 3.f, 2.f, 1.f, 2.f, 3.f, 2.f, 4.f};   // dance_moves in key order (:57-68)
    return dif[move];
  }
  template <class EN> __device__ static __forceinline__ bool upright(EN& E, int torso) {
    const float* q = E.p_xquat() + 4 * torso;
    return q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3] > 0.7f;      // element (2,2) of mju_quat2Mat
  }

  template <class EN> __device__ static void apply_action(EN& E, const TaskParams& tp, const float* act, float* act_clipped) {
    for (int i = E.lane; i < ACT; i += 32) {
      float a = clampf(act[i], tp.act_lo[i], tp.act_hi[i]);
      act_clipped[i] = a; E.p_ctrl()[i] = a;
    }
    E.sync();
  }
  // _update_rhythm + _update_visual_effects: run before mj_step, the spotlight chases the torso position of the
  // previous step's last forward pass
  template <class EN> __device__ static void pre_physics(EN& E, const TaskParams&, int* ti, float* tf) {
    if (E.lane == 0) {
      double t = D(tf, 2) + DT;
      if (t >= BEAT) { t -= BEAT; ti[1] += 1; if ((ti[1] & 3) == 0) ti[2] += 1; }
      D(tf, 2) = t;
      tf[8] += 0.1f * (tf[13] - tf[8]); tf[9] += 0.1f * (tf[14] - tf[9]); tf[10] += 0.1f * (5.0f - tf[10]);
    }
    E.sync();
  }

  template <class EN> __device__ static void reset_state(EN& E, const TaskParams& tp, const BatchView& B, int env, int* ti, float* tf,
                                     const float* inject) {
    E.reset_data();
    if (E.lane == 0) {
      float* q = E.p_qpos();
      q[2] = 1.8f; q[3] = 1.0f;      // "root height" / "quaternion w" land on abdomen_z and neck_x
      unsigned ep = (unsigned)ti[4];
      if (ep == 0) { tf[8] = 0.f; tf[9] = 0.f; tf[10] = 5.f; ti[5] = 0; }    // constructor state (:99)
      int packed[3] = {0, 0, 0};
      for (int k = 0; k < NSEQ; k++) {
        int mv; float dur;
        if (inject) { mv = (int)inject[2 * k]; dur = inject[2 * k + 1]; }
        else {
          mv = min(9, (int)(10.0f * rng_uniform(B.seed, (unsigned)(B.env_offset + env), ep, 2 * k)));
          dur = 1.0f + 2.0f * rng_uniform(B.seed, (unsigned)(B.env_offset + env), ep, 2 * k + 1);
        }
        packed[k >> 3] |= (mv & 15) << (4 * (k & 7));
        tf[16 + k] = dur;
      }
      ti[0] = 0; ti[1] = 0; ti[2] = 0; ti[3] = 0; ti[4] = (int)(ep + 1); ti[6] = 0; ti[7] = 0; ti[8] = 0;
      ti[9] = packed[0]; ti[10] = packed[1]; ti[11] = packed[2];
      tf[0] = 0.f; tf[1] = 0.5f; D(tf, 2) = 0.0; D(tf, 4) = 1.0; D(tf, 6) = 0.0; tf[11] = 0.f; tf[12] = 0.f;
      *E.p_time() = 0.f;
    }
    E.sync();
  }
  // end of reset(): prev_joint_vel <- qvel[6:], move_history <- [] (:822-826)
  template <class EN> __device__ static void after_settle(EN& E, const TaskParams& tp, int* ti, float* tf) {
    for (int i = E.lane; i < 23; i += 32) tf[36 + i] = E.p_qvel()[6 + i];
    if (E.lane < 3) tf[13 + E.lane] = E.p_xpos()[3 * tp.ids[0] + E.lane];
    if (E.lane == 0) { ti[7] = 0; ti[8] = 0; }
    E.sync();
  }

  template <class EN> __device__ static void observe(EN& E, const TaskParams& tp, float* obs) {
    int torso = tp.ids[0];
    const int* ti = E.p_ti(); float* tf = E.p_tf();
    const float* jr = E.F(DF_jnt_range);
    for (int i = E.lane; i < OBS; i += 32) {
      float v = 0.f;
      if (i < 29) {                       // qpos[7+i] normalised by the range of joint i (:1034-1049)
        if (i < 22) { float lo = jr[2 * i], hi = jr[2 * i + 1]; if (lo < hi) v = clampf(2.f * (E.p_qpos()[7 + i] - lo) / (hi - lo) - 1.f, -1.f, 1.f); }
      } else if (i < 58) { int k = i - 29; if (k < 23) v = clampf(E.p_qvel()[6 + k] / 10.0f, -1.f, 1.f); }
      else if (i < 62) v = E.p_xquat()[4 * torso + i - 58];
      else if (i < 65) v = clampf(E.p_qvel()[i - 62] / 5.0f, -1.f, 1.f);
      else if (i < 68) v = clampf(E.p_qvel()[i - 62] / 10.0f, -1.f, 1.f);
      else if (i < 71) v = clampf(E.p_rootcom()[3 * E.I(DI_body_rootidx)[torso] + i - 68] / 10.0f, -1.f, 1.f);
      else if (i < 73) {
        int foot = tp.ids[1 + i - 71], ncon = E.p_misc()[MISC_NCON];
        const int* pc1 = E.PI(DI_pair_cg1); const int* pc2 = E.PI(DI_pair_cg2); const int* gid = E.I(DI_cg_geomid);
        for (int c = 0; c < ncon; c++) {
          int p = __float_as_int(E.x_con()[B2_CON_STRIDE * c + 13]); int g1 = gid[pc1[p]], g2 = gid[pc2[p]];
          if ((g1 == foot && (g2 == tp.ids[3] || g2 == tp.ids[4])) || (g2 == foot && (g1 == tp.ids[3] || g1 == tp.ids[4]))) v = 1.f;
        }
      } else if (i < 76) v = 0.f;
      else if (i == 76) v = (float)(D(tf, 2) / BEAT);
      else if (i == 77) v = (float)((BEAT - D(tf, 2)) / BEAT);
      else if (i < 88) { if (ti[3] < NSEQ && seq_move(ti, ti[3]) == i - 78) v = 1.f; }
      else if (i == 88) v = clampf((float)(D(tf, 4) / 10.0), 0.f, 1.f);
      else if (i == 89) v = tf[1];
      else if (i < 93) v = clampf((tf[8 + i - 90] - E.p_xpos()[3 * torso + i - 90]) / 10.0f, -1.f, 1.f);
      else v = 1.0f - fminf(tf[11] / 1000.0f, 1.0f);
      obs[i] = v;
    }
  }

  template <class EN> __device__ static void post_physics(EN&, const TaskParams&, int*, float*) {}

  // lane 0: current_step += 1 -> reward -> terminated -> truncated -> stats -> crowd -> move transition -> prev state
  template <class EN> __device__ static float reward_and_done(EN& E, const TaskParams& tp, const float* act, int* ti, float* tf,
                                          int* terminated, int* truncated) {
    int torso = tp.ids[0];
    ti[0] += 1;
    const float* qv = E.p_qvel() + 6; const float* prev = tf + 36;
    double tslb = D(tf, 2), combo = D(tf, 4);
    double phase = tslb / BEAT;
    float mv2 = 0.f, dv2 = 0.f;
    for (int i = 0; i < 23; i++) { float v = qv[i], dvi = v - prev[i]; mv2 = fmaf(v, v, mv2); dv2 = fmaf(dvi, dvi, dv2); }
    float mv = sqrtf(mv2), jerk = sqrtf(dv2);
    bool up = upright(E, torso);
    float reward = 0.f;
    if (phase < 0.1 || phase > 0.9) {
      if (mv > 1.0f) { reward += 100.0f; combo = fmin(combo + 0.1, 10.0); }
      else combo = fmax(combo - 0.05, 1.0);
    }
    if (up) { reward += 30.0f; if (jerk > 0.5f) reward += 15.0f; }
    reward += 20.0f * expf(-0.1f * jerk);
    if (ti[7] > 2) { int h = ti[8], a = h & 15, b = (h >> 4) & 15, c = (h >> 8) & 15; if (a != b && b != c && a != c) reward += 50.0f; }
    double now = (double)ti[0] * DT, elapsed = now - D(tf, 6);
    int idx = ti[3];
    if (idx < NSEQ && elapsed > (double)tf[16 + idx] * 0.8) reward += 200.0f * difficulty(seq_move(ti, idx));
    {
      const float* jr = E.F(DF_jnt_range); float used = 0.f;
      for (int i = 0; i < 22; i++) { float lo = jr[2 * i], hi = jr[2 * i + 1]; if (lo < hi) used += fabsf(E.p_qpos()[7 + i] - 0.5f * (lo + hi)) / (hi - lo); }
      if (used > 5.0f) reward += 10.0f;
    }
    float a2 = 0.f, a1 = 0.f;
    for (int i = 0; i < ACT; i++) { a2 = fmaf(act[i], act[i], a2); a1 += fabsf(act[i]); }
    reward += -0.05f * a2;
    if (!up) { reward += -500.0f; combo = 1.0; }
    if (phase > 0.2 && phase < 0.8 && mv > 3.0f) reward += -5.0f;
    if (reward > 0.f) reward *= (float)combo;
    tf[0] += reward;
    // _check_termination
    int term = 0;
    if (!up) { if (ti[5] == 0) ti[5] = ti[0] + 1; else if (ti[0] - (ti[5] - 1) > 120) term = 1; }
    else ti[5] = 0;
    const float* xp = E.p_xpos() + 3 * torso;
    if (sqrtf(xp[0] * xp[0] + xp[1] * xp[1]) > 15.0f || xp[2] < 0.f || xp[2] > 5.0f) term = 1;
    *terminated = term; *truncated = ti[0] >= MAX_STEPS;
    // _update_episode_stats
    tf[11] += a1 * (float)DT;
    if (phase < 0.1 || phase > 0.9) tf[12] += (float)DT;
    ti[6] = max(ti[6], (int)combo);
    // _update_crowd_excitement
    {
      float onbeat = (tslb < 0.1 || tslb > BEAT - 0.1) ? 0.1f : 0.f;
      float cf = fminf((float)(combo / 10.0), 1.0f) * 0.2f;
      float df = idx < NSEQ ? difficulty(seq_move(ti, idx)) / 4.0f * 0.1f : 0.f;
      tf[1] = clampf(tf[1] + (onbeat + cf + df) * 0.01f, 0.f, 1.f) * 0.999f;
    }
    // _check_move_transition
    if (idx < NSEQ && elapsed >= (double)tf[16 + idx]) {
      idx += 1; ti[3] = idx; D(tf, 6) = now;
      if (idx < NSEQ) { ti[8] = ((ti[8] << 4) | seq_move(ti, idx)) & 0xfff; ti[7] += 1; }
    }
    D(tf, 4) = combo;
    for (int i = 0; i < 23; i++) tf[36 + i] = qv[i];
    tf[13] = xp[0]; tf[14] = xp[1]; tf[15] = xp[2];
    return reward;
  }
};

// ------------------------------------------------------------------------------------------------ humanoid soccer
// humanoid_soccer_env/soccer_env.py: step :398-452, reset :347-396, _randomize_initial_state :454-496,
// _update_environmental_factors :498-508, _update_goalkeeper :506-524, _apply_environmental_effects :526-537,
// _get_observation :539-631, _calculate_reward :633-690, _check_termination :692-716, _update_episode_stats :718-730,
// helpers :733-833.  Quirks kept: the robot pose is written through jnt_qposadr[0] (the goalkeeper /
// ball coordinates, so the robot itself always starts at the origin), the ball quaternion is left un-normalised, the
// goalkeeper force persists in qfrc_applied, the wind force accumulates in xfrc_applied[ball] until the next reset.
// ti: [0] current_step [1] goal_scored [2] goals_scored [3] ball_contacts [4] episode id
// tf: [0] return [1..3] prev_ball_pos [4..6] prev_robot_pos [7,8] wind force per airborne step [9,10] xfrc_applied[ball].xy
//     [11] time_upright [12] distance_traveled [13] max_ball_speed
// ids: [0] torso body [1] ball body [2] goalkeeper body [3] ball geom [4] right_foot geom [5] left_foot geom
//      [6,7] bitmask of geoms whose name contains foot/shin/thigh/torso/head/hand/arm [8] first observed joint (abdomen_y)
//      [9] goalkeeper_y joint [10] ball_joint [11] first body of the torso subtree [12] bodies in it
// inject: robot_x, robot_y, angle, 29 joint noises, goalkeeper_y, wind_strength, wind_angle, friction variation (unused)
struct SoccerTask {
  static constexpr int OBS = 80, ACT = 33, FRAME_SKIP = 1, SETTLE = 10, MAX_STEPS = 5000, NTI = 8, NTF = 16, NINJ = 36, KEEP_FRAMES = 2, XFRC_SLOT = 1, COOP_MIN = 32, ARENA_ROWS = 84, CON_CAP = 32, ARENA_SPAN = 0, MAX_EPB = 3, EPISODE_SLOT = 4, LOCKSTEP = 0, NEWTON_TEAM_ND = 16, ARENA_FLOATS = 0;
  static constexpr int SOLVER = 0;      // compile-time copy of the model's <option solver>; b2_batch_create checks it
  static constexpr bool CONDIM6 = false, RESET_FORWARD = false, PGS_HOIST = false, COLD_PAIRS = false, DYN_ISLANDS = true, CONVEX_PAIRS = false;
  static constexpr int NJOINT = 29, NOBSJ = 25;

  template <class EN> __device__ static __forceinline__ bool upright(EN& E, int torso) {
    const float* q = E.p_xquat() + 4 * torso;
    return q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3] > 0.7f;
  }
  template <class EN> __device__ static void apply_action(EN& E, const TaskParams& tp, const float* act, float* act_clipped) {
    for (int i = E.lane; i < ACT; i += 32) {
      float a = clampf(act[i], tp.act_lo[i], tp.act_hi[i]);
      act_clipped[i] = a; E.p_ctrl()[i] = a;          // mj_step clamps the four root motors to their own ctrlrange
    }
    E.sync();
  }
  // _update_goalkeeper + _apply_environmental_effects: both read the ball position of the previous forward pass,
  // which is exactly prev_ball_pos
  template <class EN> __device__ static void pre_physics(EN& E, const TaskParams& tp, int* ti, float* tf) {
    if (E.lane == 0) {
      int gq = E.I(DI_jnt_qposadr)[tp.ids[9]];
      if (tf[1] < -10.0f) {
        float err = clampf(tf[2], -3.0f, 3.0f) - E.p_qpos()[gq];
        E.p_qapp()[tp.ids[9]] = clampf(50.0f * err, -100.0f, 100.0f);      // joint id used as a dof index (:524)
      }
      if (tf[3] > 0.5f) { tf[9] += tf[7]; tf[10] += tf[8]; }
      float* x = E.p_xfrc(); x[0] = tf[9]; x[1] = tf[10]; x[2] = x[3] = x[4] = x[5] = 0.f;
    }
    E.sync();
  }
  template <class EN> __device__ static void reset_state(EN& E, const TaskParams& tp, const BatchView& B, int env, int* ti, float* tf,
                                     const float* inject) {
    E.reset_data();
    const int* jq = E.I(DI_jnt_qposadr); const float* jr = E.F(DF_jnt_range);
    unsigned ep = (unsigned)ti[4]; unsigned ge = (unsigned)(B.env_offset + env);
    auto draw = [&](int k, float lo, float hi) { return inject ? inject[k] : lo + (hi - lo) * rng_uniform(B.seed, ge, ep, (unsigned)k); };
    float* q = E.p_qpos();
    if (E.lane == 0) {
      float rx = draw(0, -15.f, -5.f), ry = draw(1, -10.f, 10.f), ang = draw(2, -0.5f, 0.5f);
      int a0 = jq[0];
      q[a0] = rx; q[a0 + 1] = ry; q[a0 + 2] = 1.4f;
      float sn, cs; sincosf(0.5f * ang, &sn, &cs);
      q[a0 + 3] = cs; q[a0 + 4] = 0.f; q[a0 + 5] = 0.f; q[a0 + 6] = sn;
      int bq = jq[tp.ids[10]];
      q[bq] = rx + 2.0f; q[bq + 1] = ry; q[bq + 2] = 0.15f;
    }
    E.sync();
    for (int i = E.lane; i < NJOINT; i += 32) {
      int j = tp.ids[8] + i; float lo = jr[2 * j], hi = jr[2 * j + 1];
      float noise = draw(3 + i, -0.1f, 0.1f);
      if (lo < hi) q[jq[j]] = clampf(0.5f * (lo + hi) + noise, lo, hi);
    }
    E.sync();
    if (E.lane == 0) {
      q[jq[tp.ids[9]]] = draw(32, -2.f, 2.f);
      float ws = draw(33, 0.f, 2.f), wa = draw(34, 0.f, 6.283185307179586f);
      float sn, cs; sincosf(wa, &sn, &cs);
      tf[7] = ws * cs * 0.1f; tf[8] = ws * sn * 0.1f; tf[9] = 0.f; tf[10] = 0.f;
      tf[0] = 0.f; tf[11] = 0.f; tf[12] = 0.f; tf[13] = 0.f;
      ti[0] = 0; ti[1] = 0; ti[2] = 0; ti[3] = 0; ti[4] = (int)(ep + 1);
      float* x = E.p_xfrc(); for (int k = 0; k < 6; k++) x[k] = 0.f;
      *E.p_time() = 0.f;
    }
    E.sync();
  }
  template <class EN> __device__ static void after_settle(EN& E, const TaskParams& tp, int* ti, float* tf) {
    if (E.lane < 3) { tf[1 + E.lane] = E.p_xpos()[3 * tp.ids[1] + E.lane]; tf[4 + E.lane] = E.p_xpos()[3 * tp.ids[0] + E.lane]; }
    E.sync();
  }
  // current_step += 1 precedes the observation (:416-419), which reads it (time remaining)
  template <class EN> __device__ static void post_physics(EN& E, const TaskParams&, int* ti, float*) { if (E.lane == 0) ti[0] += 1; E.sync(); }

  template <class EN> __device__ static __forceinline__ bool robot_geom(const TaskParams& tp, int g) {
    return g < 64 && ((g < 32 ? (unsigned)tp.ids[6] >> g : (unsigned)tp.ids[7] >> (g - 32)) & 1u);
  }
  template <class EN> __device__ static void observe(EN& E, const TaskParams& tp, float* obs) {
    const int torso = tp.ids[0], ball = tp.ids[1], gk = tp.ids[2];
    const int* ti = E.p_ti();
    const int* jq = E.I(DI_jnt_qposadr); const int* jd = E.I(DI_jnt_dofadr); const float* jr = E.F(DF_jnt_range);
    const float* xp = E.p_xpos();
    for (int i = E.lane; i < OBS; i += 32) {
      float v = 0.f;
      if (i < 25) { int j = tp.ids[8] + i; float lo = jr[2 * j], hi = jr[2 * j + 1]; if (lo < hi) v = clampf(2.f * (E.p_qpos()[jq[j]] - lo) / (hi - lo) - 1.f, -1.f, 1.f); }
      else if (i < 50) v = clampf(E.p_qvel()[jd[tp.ids[8] + i - 25]] / 10.0f, -1.f, 1.f);
      else if (i < 54) v = E.p_xquat()[4 * torso + i - 50];
      else if (i < 57) v = clampf(E.p_qvel()[i - 54] / 5.0f, -1.f, 1.f);
      else if (i < 60) v = clampf(E.p_qvel()[i - 54] / 10.0f, -1.f, 1.f);
      else if (i < 63) v = clampf((xp[3 * ball + i - 60] - xp[3 * torso + i - 60]) / 30.0f, -1.f, 1.f);
      else if (i < 66) v = clampf(E.p_qvel()[jd[tp.ids[10]] + i - 63] / 20.0f, -1.f, 1.f);
      else if (i < 69) { const float g[3] = {24.5f, 0.0f, 1.22f}; v = clampf((g[i - 66] - xp[3 * torso + i - 66]) / 30.0f, -1.f, 1.f); }
      else if (i < 73) {
        // (dist, |friction[:2]|) of the last contact between a foot geom and geom 0 (:765-785)
        int foot = tp.ids[4 + ((i - 69) >> 1)], ncon = E.p_misc()[MISC_NCON];
        const int* pc1 = E.PI(DI_pair_cg1); const int* pc2 = E.PI(DI_pair_cg2); const int* gid = E.I(DI_cg_geomid);
        const int* pprm = E.PI(DI_pair_prm); const float* prm = E.F(DF_prm);
        for (int c = 0; c < ncon; c++) {
          int p = __float_as_int(E.x_con()[B2_CON_STRIDE * c + 13]); int g1 = gid[pc1[p]], g2 = gid[pc2[p]];
          if ((g1 == foot && g2 == 0) || (g2 == foot && g1 == 0)) {
            const float* pr = prm + B2DEV_PRM_STRIDE * pprm[p];
            v = ((i - 69) & 1) ? sqrtf(pr[2] * pr[2] + pr[3] * pr[3]) : E.x_con()[B2_CON_STRIDE * c];
          }
        }
        v = clampf(v / 1000.0f, -1.f, 1.f);
      } else if (i < 76) {
        // subtree_com[torso]: mass-weighted com of the bodies below the torso (the `root` body is not part of it)
        const float* mass = E.F(DF_body_mass); float s = 0.f, m = 0.f;
        for (int b = tp.ids[11]; b < tp.ids[11] + tp.ids[12]; b++) { s = fmaf(mass[b], E.p_xipos()[3 * b + i - 73], s); m += mass[b]; }
        v = clampf(s / m / 30.0f, -1.f, 1.f);
      } else if (i == 76) v = 1.0f - (float)ti[0] / (float)MAX_STEPS;
      else if (i == 77) {
        float dx = xp[3 * ball] - xp[3 * torso], dy = xp[3 * ball + 1] - xp[3 * torso + 1], dz = xp[3 * ball + 2] - xp[3 * torso + 2];
        v = clampf(sqrtf(dx * dx + dy * dy + dz * dz) / 50.0f, 0.f, 1.f);
      } else v = clampf(xp[3 * gk + i - 78] / 15.0f, -1.f, 1.f);
      obs[i] = v;
    }
  }

  template <class EN> __device__ static float reward_and_done(EN& E, const TaskParams& tp, const float* act, int* ti, float* tf,
                                          int* terminated, int* truncated) {
    const int torso = tp.ids[0], ball = tp.ids[1];
    V3 bp = ld3(E.p_xpos() + 3 * ball), rp = ld3(E.p_xpos() + 3 * torso);
    V3 pb = ld3(tf + 1), pr = ld3(tf + 4);
    float reward = 0.f;
    if (bp.x > 24.0f && fabsf(bp.y) < 3.66f && bp.z < 2.44f) { reward += 10000.0f; ti[1] = 1; ti[2] += 1; }
    {
      int ncon = E.p_misc()[MISC_NCON]; bool touch = false;
      const int* pc1 = E.PI(DI_pair_cg1); const int* pc2 = E.PI(DI_pair_cg2); const int* gid = E.I(DI_cg_geomid);
      for (int c = 0; c < ncon && !touch; c++) {
        int p = __float_as_int(E.x_con()[B2_CON_STRIDE * c + 13]); int g1 = gid[pc1[p]], g2 = gid[pc2[p]];
        if (g1 == tp.ids[3] || g2 == tp.ids[3]) touch = robot_geom<EN>(tp, g1 == tp.ids[3] ? g2 : g1);
      }
      if (touch) { reward += 1000.0f; ti[3] += 1; }
    }
    float cur = norm(bp - rp), prev = norm(pb - pr);
    if (cur < prev && cur > 2.0f) reward += 500.0f * (prev - cur);
    bool up = upright(E, torso);
    if (up) { reward += 200.0f; tf[11] += 0.02f; }
    V3 goal = v3(24.5f, 0.f, 0.f);
    float pg = norm(pr - goal), cg = norm(rp - goal);
    if (cg < pg) reward += 100.0f * (pg - cg);
    float a2 = 0.f;
    for (int i = 0; i < ACT; i++) a2 = fmaf(act[i], act[i], a2);
    reward += -0.1f * a2;
    if (!up) reward += -1000.0f;
    float pbg = norm(pb - goal), cbg = norm(bp - goal);
    if (cbg < pbg) reward += 300.0f * (pbg - cbg);
    tf[0] += reward;
    *terminated = ti[1] || (!up && ti[0] > 100) || fabsf(bp.x) > 30.0f || fabsf(bp.y) > 20.0f || bp.z < -1.0f || bp.z > 10.0f ||
                  fabsf(rp.x) > 30.0f || fabsf(rp.y) > 20.0f || rp.z < 0.0f || rp.z > 5.0f;
    *truncated = ti[0] >= MAX_STEPS;
    tf[12] += norm(rp - pr);
    { const float* bv = E.p_qvel() + E.I(DI_jnt_dofadr)[tp.ids[10]]; tf[13] = fmaxf(tf[13], sqrtf(bv[0] * bv[0] + bv[1] * bv[1] + bv[2] * bv[2])); }
    st3(tf + 1, bp); st3(tf + 4, rp);
    return reward;
  }
};

// ------------------------------------------------------------------------------------------------ bipedal rescue
// bipedal_rescue_env/rescue_env.py: step :416-471, reset :366-414, _randomize_initial_state :473-508,
// _check_victim_interactions :510-543, _get_observation :545-600, _calculate_reward :602-668, _check_termination
// :670-697, _update_episode_stats :699-706, helpers :708-775.  Quirks kept: the carry-capacity test
// is made once before the victim loop (more than two victims can be picked up in one step), the first step of every
// episode pays +inf (approach term against closest = inf), the attributes created with hasattr
// (_prev_rescued_count, _prev_carried_count, _prev_safe_zone_distance, _fall_timer) survive reset.
// ti: [0] current_step [1] rescued mask [2] carried mask [3] carrying flag [4] episode id [5] _prev_rescued_count + 1
//     (0 = attribute absent) [6] _prev_carried_count + 1 [7] _fall_timer [8] has _prev_safe_zone_distance [9] falls
//     [10] collisions [11] victims_rescued
// tf: [0] return [1] current_energy [2] closest_victim_distance [3] _prev_safe_zone_distance [4,5] prev_robot_pos.xy
//     [6] distance_traveled [7] energy_used [8] time_to_first_rescue (-1 = None)
// ids: [0] torso body [1] victim1 body (victims are consecutive) [2] root_x joint (y, z follow) [3] first observed joint
//      (26 consecutive) [4] victim1_x joint (victim joints are 6 apart, y = x + 1)
// inject: robot_x, robot_y, then (x_offset, y_offset) for the five victims
struct RescueTask {
  static constexpr int OBS = 102, ACT = 26, FRAME_SKIP = 1, SETTLE = 10, MAX_STEPS = 10000, NTI = 12, NTF = 12, NINJ = 12, KEEP_FRAMES = 1, XFRC_SLOT = -1, COOP_MIN = 32, ARENA_ROWS = 92, CON_CAP = 48, ARENA_SPAN = 40, MAX_EPB = 3, EPISODE_SLOT = 4, LOCKSTEP = 0, NEWTON_TEAM_ND = 16, ARENA_FLOATS = 7400;
  static constexpr int SOLVER = 0;      // compile-time copy of the model's <option solver>; b2_batch_create checks it
  static constexpr bool CONDIM6 = false, RESET_FORWARD = false, PGS_HOIST = false, COLD_PAIRS = true, DYN_ISLANDS = true, CONVEX_PAIRS = true;
  static constexpr int NVICT = 5;

  template <class EN> __device__ static __forceinline__ bool upright(EN& E, int torso) {
    const float* q = E.p_xquat() + 4 * torso;
    return q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3] > 0.7f;
  }
  template <class EN> __device__ static void apply_action(EN& E, const TaskParams& tp, const float* act, float* act_clipped) {
    for (int i = E.lane; i < ACT; i += 32) {
      float a = clampf(act[i], tp.act_lo[i], tp.act_hi[i]);
      act_clipped[i] = a; E.p_ctrl()[i] = a;
    }
    E.sync();
  }
  // energy bookkeeping happens before mj_step (:424-427)
  template <class EN> __device__ static void pre_physics(EN& E, const TaskParams&, int*, float* tf) {
    if (E.lane == 0) {
      const float* a = E.p_act(); float s = 0.f;
      for (int i = 0; i < ACT; i++) s += fabsf(a[i]);
      tf[1] -= s * 0.001f; tf[7] += s * 0.001f;
    }
    E.sync();
  }
  template <class EN> __device__ static void reset_state(EN& E, const TaskParams& tp, const BatchView& B, int env, int* ti, float* tf,
                                     const float* inject) {
    E.reset_data();
    if (E.lane == 0) {
      const int* jq = E.I(DI_jnt_qposadr); float* q = E.p_qpos();
      unsigned ep = (unsigned)ti[4]; unsigned ge = (unsigned)(B.env_offset + env);
      auto draw = [&](int k, float lo, float hi) { return inject ? inject[k] : lo + (hi - lo) * rng_uniform(B.seed, ge, ep, (unsigned)k); };
      q[jq[tp.ids[2]]] = draw(0, -5.f, 5.f); q[jq[tp.ids[2] + 1]] = draw(1, -5.f, 5.f); q[jq[tp.ids[2] + 2]] = 1.2f;
      for (int v = 0; v < NVICT; v++) {
        int jx = tp.ids[4] + 6 * v;
        q[jq[jx]] += draw(2 + 2 * v, -1.f, 1.f); q[jq[jx + 1]] += draw(3 + 2 * v, -1.f, 1.f);
      }
      ti[0] = 0; ti[1] = 0; ti[2] = 0; ti[3] = 0; ti[4] = (int)(ep + 1); ti[9] = 0; ti[10] = 0; ti[11] = 0;
      tf[0] = 0.f; tf[1] = 1000.0f; tf[2] = __int_as_float(0x7f800000); tf[6] = 0.f; tf[7] = 0.f; tf[8] = -1.f;
      *E.p_time() = 0.f;
    }
    E.sync();
  }
  template <class EN> __device__ static void after_settle(EN& E, const TaskParams& tp, int*, float* tf) {
    if (E.lane < 2) tf[4 + E.lane] = E.p_xpos()[3 * tp.ids[0] + E.lane];
    E.sync();
  }
  // current_step += 1, then _check_victim_interactions, both before the observation (:432-438)
  template <class EN> __device__ static void post_physics(EN& E, const TaskParams& tp, int* ti, float* tf) {
    if (E.lane == 0) {
      ti[0] += 1;
      const float* xp = E.p_xpos(); float rx = xp[3 * tp.ids[0]], ry = xp[3 * tp.ids[0] + 1];
      if (__popc(ti[2]) < 2) {                       // capacity is tested once, before the loop
        for (int v = 0; v < NVICT; v++) {
          if (((ti[1] | ti[2]) >> v) & 1) continue;
          float dx = rx - xp[3 * (tp.ids[1] + v)], dy = ry - xp[3 * (tp.ids[1] + v) + 1];
          float d = sqrtf(dx * dx + dy * dy);
          if (d < 1.0f && d < 0.8f) { ti[2] |= 1 << v; ti[3] = 1; }
        }
      }
      if (ti[3]) {
        float dx = rx - 20.0f, dy = ry;
        if (sqrtf(dx * dx + dy * dy) < 3.0f) {
          // victims_rescued is a list: len() counts every drop-off
          int n = __popc(ti[2]);
          if (n > 0 && tf[8] < 0.f) tf[8] = (float)ti[0] * 0.02f;
          ti[11] += n; ti[1] |= ti[2]; ti[2] = 0; ti[3] = 0;
        }
      }
    }
    E.sync();
  }
  template <class EN> __device__ static void observe(EN& E, const TaskParams& tp, float* obs) {
    const int torso = tp.ids[0];
    const int* ti = E.p_ti(); const float* tf = E.p_tf();
    const int* jq = E.I(DI_jnt_qposadr); const int* jd = E.I(DI_jnt_dofadr);
    const float* xp = E.p_xpos();
    for (int i = E.lane; i < OBS; i += 32) {
      float v = 0.f;
      if (i < 52) { int j = tp.ids[3] + (i >> 1); v = (i & 1) ? E.p_qvel()[jd[j]] : E.p_qpos()[jq[j]]; }
      else if (i < 55) v = xp[3 * torso + i - 52];
      else if (i < 59) v = E.p_xquat()[4 * torso + i - 55];
      else if (i < 65) v = E.p_qvel()[jd[tp.ids[2]] + i - 59];
      else if (i == 65) { int n = min(E.p_misc()[MISC_NCON], 10); for (int c = 0; c < n; c++) v += fabsf(E.x_con()[B2_CON_STRIDE * c]); }
      else if (i < 69) v = 0.f;
      else if (i < 89) {
        int k = (i - 69) >> 2, f = (i - 69) & 3;
        v = f < 2 ? xp[3 * (tp.ids[1] + k) + f] : (f == 2 ? (float)((ti[1] >> k) & 1) : (float)((ti[2] >> k) & 1));
      } else if (i < 92) { const float sz[3] = {20.0f, 0.0f, 0.0f}; v = sz[i - 89] - xp[3 * torso + i - 89]; }
      else if (i == 92) v = tf[1] / 1000.0f;
      else if (i == 93) v = 1.0f - (float)ti[0] / (float)MAX_STEPS;
      else if (i == 94) v = (float)__popc(ti[2]);
      else if (i == 95) v = (float)ti[11];
      else { const float fz[6] = {-5.0f, -3.0f, 0.0f, 8.0f, 6.0f, 0.0f}; v = fz[i - 96] - xp[3 * torso + (i - 96) % 3]; }
      obs[i] = v;
    }
  }
  template <class EN> __device__ static float reward_and_done(EN& E, const TaskParams& tp, const float* act, int* ti, float* tf,
                                          int* terminated, int* truncated) {
    const int torso = tp.ids[0];
    const float* xp = E.p_xpos(); float rx = xp[3 * torso], ry = xp[3 * torso + 1];
    float reward = 0.f;
    int nres = ti[11], ncar = __popc(ti[2]);
    if (ti[5] > 0) { int d = nres - (ti[5] - 1); if (d > 0) reward += 5000.0f * (float)d; }
    ti[5] = nres + 1;
    if (ti[6] > 0) { int d = ncar - (ti[6] - 1); if (d > 0) reward += 1000.0f * (float)d; }
    ti[6] = ncar + 1;
    float mind = __int_as_float(0x7f800000);
    for (int v = 0; v < NVICT; v++) {
      if (((ti[1] | ti[2]) >> v) & 1) continue;
      float dx = rx - xp[3 * (tp.ids[1] + v)], dy = ry - xp[3 * (tp.ids[1] + v) + 1];
      mind = fminf(mind, sqrtf(dx * dx + dy * dy));
    }
    if (mind < tf[2] && mind < 10.0f) reward += 100.0f * (tf[2] - mind);       // +inf on the first step of an episode
    tf[2] = mind;
    if (ti[3]) {
      float dx = rx - 20.0f, dy = ry, sd = sqrtf(dx * dx + dy * dy);
      if (ti[8] && sd < tf[3]) reward += 200.0f * (tf[3] - sd);
      tf[3] = sd; ti[8] = 1;
    }
    bool up = upright(E, torso);
    if (up) reward += 50.0f; else { reward += -500.0f; ti[9] += 1; }
    float s = 0.f;
    for (int i = 0; i < ACT; i++) s += fabsf(act[i]);
    if (s * 0.001f < 0.5f) reward += 10.0f;
    { float dx = rx + 5.0f, dy = ry + 3.0f; if (sqrtf(dx * dx + dy * dy) < 1.5f) reward += -200.0f; }
    { float dx = rx - 8.0f, dy = ry - 6.0f; if (sqrtf(dx * dx + dy * dy) < 1.2f) reward += -200.0f; }
    {
      int n = min(E.p_misc()[MISC_NCON], 20); bool hit = false;
      for (int c = 0; c < n; c++) if (fabsf(E.x_con()[B2_CON_STRIDE * c]) > 0.1f) hit = true;
      if (hit) { reward += -100.0f; ti[10] += 1; }
    }
    reward += -1.0f;
    tf[0] += reward;
    int term = (ti[11] == NVICT);
    // the reference returns at the first true test: the fall timer only moves when "all rescued" was false
    if (!term) { if (!up) { ti[7] += 1; if (ti[7] > 100) term = 1; } else ti[7] = 0; }
    if (tf[1] <= 0.f) term = 1;
    if (fabsf(rx) > 25.0f || fabsf(ry) > 25.0f) term = 1;
    *terminated = term; *truncated = ti[0] >= MAX_STEPS;
    { float dx = rx - tf[4], dy = ry - tf[5]; tf[6] += sqrtf(dx * dx + dy * dy); }
    tf[4] = rx; tf[5] = ry;
    return reward;
  }
};

// ------------------------------------------------------------------------------------------------ humanoid construction
// humanoid_construction_env/construction_env.py: reset :547-584, step :586-623, _get_observation :625-659,
// _calculate_reward :661-700, _update_task_progress :702-719, _check_terminated :721-737.  RK4 at 2 ms,
// MuJoCo's default solver (Newton, 100 iterations, 1e-8).  The observation has 135 entries (the reference declares 125,
// SURVEY F11) and is taken after reward / termination; reset neither settles nor calls mj_forward.
// ti: [0] current_step [1] current_task (0 stack_blocks, 1 operate_crane, 2 transport_material, 3 build_structure)
//     [2] episode id [3] tasks_completed
// tf: [0] total_reward [1] task_progress [2] wind_strength [3] rain_intensity [4] temperature
// ids: [0] humanoid body      inject: task index, wind, rain, temperature
struct ConstructionTask {
  static constexpr int OBS = 135, ACT = 33, FRAME_SKIP = 1, SETTLE = 0, MAX_STEPS = 3000, NTI = 4, NTF = 8, NINJ = 4, KEEP_FRAMES = 0, XFRC_SLOT = -1, COOP_MIN = 32, ARENA_ROWS = 92, CON_CAP = 96, ARENA_SPAN = 40, MAX_EPB = 2, EPISODE_SLOT = 2, LOCKSTEP = 0, NEWTON_TEAM_ND = 16, ARENA_FLOATS = 14000;
  static constexpr int SOLVER = 2;      // compile-time copy of the model's <option solver>; b2_batch_create checks it
  static constexpr bool CONDIM6 = false, RESET_FORWARD = false, PGS_HOIST = false, COLD_PAIRS = true, DYN_ISLANDS = true, CONVEX_PAIRS = false;

  template <class EN> __device__ static void apply_action(EN& E, const TaskParams& tp, const float* act, float* act_clipped) {
    for (int i = E.lane; i < ACT; i += 32) {
      float a = clampf(act[i], tp.act_lo[i], tp.act_hi[i]);
      act_clipped[i] = a; E.p_ctrl()[i] = a;
    }
    E.sync();
  }
  template <class EN> __device__ static void pre_physics(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void after_settle(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void reset_state(EN& E, const TaskParams& tp, const BatchView& B, int env, int* ti, float* tf,
                                     const float* inject) {
    E.reset_data();
    if (E.lane == 0) {
      unsigned ep = (unsigned)ti[2]; unsigned ge = (unsigned)(B.env_offset + env);
      auto draw = [&](int k, float lo, float hi) { return inject ? inject[k] : lo + (hi - lo) * rng_uniform(B.seed, ge, ep, (unsigned)k); };
      ti[0] = 0; ti[1] = inject ? (int)inject[0] : min(3, (int)(4.0f * rng_uniform(B.seed, ge, ep, 0))); ti[2] = (int)(ep + 1);
      tf[0] = 0.f; tf[1] = 0.f; tf[2] = draw(1, 0.f, 5.f); tf[3] = draw(2, 0.f, 0.5f); tf[4] = draw(3, 15.f, 35.f);
      *E.p_time() = 0.f;
    }
    E.sync();
  }
  // current_step += 1 and _update_task_progress precede the reward (:597-600)
  template <class EN> __device__ static void post_physics(EN& E, const TaskParams&, int* ti, float* tf) {
    if (E.lane == 0) {
      ti[0] += 1;
      float pr = 0.f;
      if (ti[1] == 1) pr = fminf(1.0f, (float)ti[0] / 500.0f); else if (ti[1] == 2) pr = fminf(1.0f, (float)ti[0] / 300.0f);
      tf[1] = pr;
    }
    E.sync();
  }
  template <class EN> __device__ static void observe(EN& E, const TaskParams&, float* obs) {
    const int* ti = E.p_ti(); const float* tf = E.p_tf();
    for (int i = E.lane; i < OBS; i += 32) {
      float v = 0.f;
      if (i < 30) v = E.p_qpos()[i];
      else if (i < 60) v = E.p_qvel()[i - 30];
      else if (i >= 90 && i < 94) v = (ti[1] == i - 90) ? 1.f : 0.f;
      else if (i == 94) v = tf[1];
      else if (i == 100) v = tf[2] / 10.0f;
      else if (i == 101) v = tf[3];
      else if (i == 102) v = tf[4] / 50.0f;
      else if (i == 110) v = 1.0f;
      obs[i] = v;
    }
  }
  // the kernel observes before this hook; the reference observes after it, but nothing the observation reads changes here
  template <class EN> __device__ static float reward_and_done(EN& E, const TaskParams& tp, const float* act, int* ti, float* tf,
                                          int* terminated, int* truncated) {
    float reward = 0.f;
    if (ti[1] == 0) reward += tf[1] * 500.0f; else if (ti[1] == 1) reward += 20.0f; else if (ti[1] == 2) reward += 30.0f; else reward += tf[1] * 100.0f;
    reward += 1.0f;
    float s = 0.f;
    for (int i = 0; i < ACT; i++) s += fabsf(act[i]);
    reward += -0.2f * s;
    float z = E.p_xpos()[3 * tp.ids[0] + 2];
    reward += z > 1.0f ? 5.0f : -2000.0f;
    int term = 0;
    if (z < 0.5f) term = 1;
    else if (tf[1] >= 1.0f) { ti[3] += 1; term = 1; }
    *terminated = term; *truncated = ti[0] >= MAX_STEPS;
    tf[0] += reward;
    return reward;
  }
};

// ------------------------------------------------------------------------------------------------ martial arts
// humanoid_martial_arts_env/martial_arts_env.py: reset :442-487, step :489-523, _get_observation :525-560,
// _calculate_reward :562-606, _check_termination :608-621.  One Euler step of 16.67 ms, Newton-50/1e-10.
// qpos[0:7] is dummy #1's free joint: reset drops that dummy next to the humanoid.  cvel[:3] (angular) is
// what the reference calls the linear velocity and vice versa; both are the last forward pass's values.  Reset ends with
// mj_forward (RESET_FORWARD).  The observation has 113 entries (85 declared) and precedes the reward, which
// advances stance_stability_time.
// ti: [0] current_step [1] techniques_performed [2] episode id [3] falls
// tf: [0] total reward [1] stance_stability_time
// ids: [0] torso [1] right_hand [2] left_hand [3] right_ankle [4] left_ankle [5] dummy1 [6] dummy2     inject: dx, dy
#ifndef B2_MARTIAL_EPB
#define B2_MARTIAL_EPB 2      // teams per SM (A/B on B200 with the team-wide Newton solve: 2 -> 969 k, 3 -> 897 k env-steps/s)
#endif
struct MartialArtsTask {
  static constexpr int OBS = 113, ACT = 28, FRAME_SKIP = 1, SETTLE = 0, MAX_STEPS = 6000, NTI = 4, NTF = 4, NINJ = 2, KEEP_FRAMES = 2, XFRC_SLOT = -1, COOP_MIN = 32, ARENA_ROWS = 92, CON_CAP = 80, ARENA_SPAN = 47, MAX_EPB = B2_MARTIAL_EPB, EPISODE_SLOT = 2, LOCKSTEP = 0, NEWTON_TEAM_ND = 16, ARENA_FLOATS = 0;
  static constexpr int SOLVER = 2;      // compile-time copy of the model's <option solver>; b2_batch_create checks it
  static constexpr bool CONDIM6 = false, RESET_FORWARD = true, PGS_HOIST = false, COLD_PAIRS = false, DYN_ISLANDS = true, CONVEX_PAIRS = true;

  template <class EN> __device__ static void apply_action(EN& E, const TaskParams& tp, const float* act, float* act_clipped) {
    for (int i = E.lane; i < ACT; i += 32) {
      float a = clampf(act[i], -1.0f, 1.0f);
      act_clipped[i] = a; E.p_ctrl()[i] = a * tp.act_hi[i];          // action * actuator_ctrlrange[:, 1] (:495)
    }
    E.sync();
  }
  template <class EN> __device__ static void pre_physics(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void after_settle(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void reset_state(EN& E, const TaskParams&, const BatchView& B, int env, int* ti, float* tf,
                                     const float* inject) {
    E.reset_data();
    if (E.lane == 0) {
      unsigned ep = (unsigned)ti[2]; unsigned ge = (unsigned)(B.env_offset + env);
      float dx = inject ? inject[0] : -0.5f + rng_uniform(B.seed, ge, ep, 0u), dy = inject ? inject[1] : -0.5f + rng_uniform(B.seed, ge, ep, 1u);
      float* q = E.p_qpos();
      q[0] = 0.0f + dx; q[1] = 0.0f + dy; q[2] = 1.4f; q[3] = 1.f; q[4] = 0.f; q[5] = 0.f; q[6] = 0.f;
      ti[0] = 0; ti[1] = 0; ti[2] = (int)(ep + 1); ti[3] = 0; tf[0] = 0.f; tf[1] = 0.f;
      *E.p_time() = 0.f;
    }
    E.sync();
  }
  template <class EN> __device__ static void post_physics(EN& E, const TaskParams&, int* ti, float*) {
    if (E.lane == 0) ti[0] += 1;
    E.sync();
  }
  template <class EN> __device__ static void observe(EN& E, const TaskParams& tp, float* obs) {
    const float* tf = E.p_tf(); const int torso = tp.ids[0];
    for (int i = E.lane; i < OBS; i += 32) {
      float v = 0.f;
      if (i < 3) v = E.p_xpos()[3 * torso + i];
      else if (i < 7) v = E.p_xquat()[4 * torso + i - 3];
      else if (i < 13) v = E.p_cvel()[6 * torso + i - 7];
      else if (i < 56) v = E.p_qpos()[7 + i - 13];
      else if (i < 97) v = E.p_qvel()[6 + i - 56];
      else if (i < 100) v = E.p_xpos()[3 * tp.ids[5] + i - 97];
      else if (i < 103) v = E.p_xpos()[3 * tp.ids[6] + i - 100];
      else if (i == 104) v = -2.0f;
      else if (i == 105) v = 1.0f;
      else if (i == 112) v = tf[1];
      obs[i] = v;
    }
  }
  template <class EN> __device__ static float reward_and_done(EN& E, const TaskParams& tp, const float* act, int* ti, float* tf,
                                          int* terminated, int* truncated) {
    const float dt = 0.01667f;
    const float* xp = E.p_xpos(); const float* cv = E.p_cvel(); const int torso = tp.ids[0];
    auto n3 = [&](const float* v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    float reward = 100.0f * fminf(1.0f, xp[3 * torso + 2] / 1.75f);
    if (n3(cv + 6 * tp.ids[1]) > 2.0f || n3(cv + 6 * tp.ids[2]) > 2.0f) { reward += 500.f; ti[1] += 1; }
    if (n3(cv + 6 * tp.ids[3]) > 3.0f || n3(cv + 6 * tp.ids[4]) > 3.0f) { reward += 800.f; ti[1] += 1; }
    if (n3(cv + 6 * torso + 3) < 0.5f) { tf[1] += dt; reward += 200.f * dt; }
    float s = 0.f;
    for (int i = 0; i < ACT; i++) s += fabsf(act[i]);
    reward -= s * 0.01f;
    float ddx = xp[3 * tp.ids[5]] - xp[3 * torso], ddy = xp[3 * tp.ids[5] + 1] - xp[3 * torso + 1];      // active_dummy_idx stays 0
    float dist = sqrtf(ddx * ddx + ddy * ddy);
    if (dist < 2.0f) reward += 50.f * (2.0f - dist);
    int term = 0;
    if (xp[3 * torso + 2] < 0.5f) { ti[3] += 1; term = 1; }
    else if (fabsf(xp[3 * torso]) > 5.5f || fabsf(xp[3 * torso + 1]) > 5.5f) term = 1;
    *terminated = term; *truncated = ti[0] >= MAX_STEPS;
    tf[0] += reward;
    return reward;
  }
};

// ------------------------------------------------------------------------------------------------ robotic arm assembly
// robotic_arm_assembly_env/assembly_env.py: reset :162-192 (+ _reset_components :194-218, 10 settle steps), step :220-250,
// _apply_action :252-265, _update_task_state :267-297, _get_gripper_contacts :299-322, _calculate_reward :331-387,
// _get_max_contact_force :389-397, _check_termination :399-417, _get_observation :419-472.
// Ten Euler sub-steps of 2 ms, Newton-50/1e-10, 18 condim-6 pad pairs (CONDIM6).  The reference matches geom *names* by
// substring per contact; the host resolves that once into aux_i[geom] (0..8 component, 100 gripper pad, -1 neither).
// `list(set(contacts))[0]` is hash-order dependent in the reference; the first component in contact order is taken.
// ti: [0] step_count [1] assembly_progress bits [2] episode id [3] held component (-1 none) [4] task phase (0 idle 1 pickup
//     2 transport 4 insert) [5..13] component status (0 in_bin 1 held 2 assembled 3 dropped)
// tf: [0] cumulative_reward
// ids: [0..8] component bodies in assembly order, [9] body of ee_site    aux_f: [3k..3k+2] target of component k, [27..29] ee_site offset
#ifndef B2_ARM_TEAM_ND
#define B2_ARM_TEAM_ND (1 << 20)      // the arm's islands are the 9-dof arm plus the 6-dof parts it touches: one warp each (A/B below)
#endif
struct ArmTask {
  static constexpr int OBS = 110, ACT = 9, FRAME_SKIP = 10, SETTLE = 10, MAX_STEPS = 150000, NTI = 16, NTF = 4, NINJ = 1, KEEP_FRAMES = 2, XFRC_SLOT = -1, COOP_MIN = 32, ARENA_ROWS = 128, CON_CAP = 128, ARENA_SPAN = 9, MAX_EPB = 4, EPISODE_SLOT = 2, LOCKSTEP = 1, NEWTON_TEAM_ND = B2_ARM_TEAM_ND, ARENA_FLOATS = 4000;
  static constexpr int SOLVER = 2;
  static constexpr bool CONDIM6 = true, RESET_FORWARD = false, PGS_HOIST = false, COLD_PAIRS = true, DYN_ISLANDS = true, CONVEX_PAIRS = true;

  template <class EN> __device__ static void apply_action(EN& E, const TaskParams& tp, const float* act, float* act_clipped) {
    for (int i = E.lane; i < ACT; i += 32) {
      float a = clampf(act[i], tp.act_lo[i], tp.act_hi[i]);
      act_clipped[i] = a;
      if (i < 7) E.p_ctrl()[i] = a;
      else if (i == 7) { E.p_ctrl()[7] = a / 1000.0f; E.p_ctrl()[8] = a / 1000.0f; }      // opening in mm for both fingers; the force entry is unused (:261-265)
    }
    E.sync();
  }
  template <class EN> __device__ static void pre_physics(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void after_settle(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void reset_state(EN& E, const TaskParams& tp, const BatchView&, int, int* ti, float* tf,
                                     const float*) {
    E.reset_data();
    const int* jq = E.I(DI_jnt_qposadr); const int* bja = E.I(DI_body_jntadr);
    if (E.lane == 0) {
      float* q = E.p_qpos();
      q[0] = 0.f; q[1] = -0.5f; q[2] = 0.5f; q[3] = 0.f; q[4] = 0.5f; q[5] = 0.f; q[6] = 0.f;             // home (:172)
      const float ini[27] = {-0.6f, 0.3f, 0.76f, -0.6f, -0.3f, 0.76f, -0.58f, -0.3f, 0.76f, -0.62f, -0.3f, 0.76f, -0.6f, -0.28f, 0.76f,
                             -0.6f, 0.f, 0.76f, 0.6f, 0.3f, 0.76f, 0.6f, -0.3f, 0.76f, 0.6f, 0.f, 0.76f};                      // :197-207 in assembly order
      for (int k = 0; k < 9; k++) {
        int a = jq[bja[tp.ids[k]]];
        q[a] = ini[3 * k]; q[a + 1] = ini[3 * k + 1]; q[a + 2] = ini[3 * k + 2]; q[a + 3] = 1.f; q[a + 4] = 0.f; q[a + 5] = 0.f; q[a + 6] = 0.f;
      }
      int ep = ti[2];
      for (int k = 0; k < NTI; k++) ti[k] = 0;
      ti[2] = ep + 1; ti[3] = -1; tf[0] = 0.f;
      *E.p_time() = 0.f;
    }
    E.sync();
  }
  template <class EN> __device__ static float max_contact_force(EN& E) {
    float mx = 0.f; const int ncon = E.p_misc()[MISC_NCON];
    for (int c = 0; c < ncon; c++) mx = fmaxf(mx, fabsf(E.x_con()[B2_CON_STRIDE * c]) * 1000.0f);
    return mx;
  }
  // step_count += 1 (:222) and _update_task_state (:267-297)
  template <class EN> __device__ static void post_physics(EN& E, const TaskParams& tp, int* ti, float*) {
    if (E.lane == 0) {
      ti[0] += 1;
      const int* pc1 = E.PI(DI_pair_cg1); const int* pc2 = E.PI(DI_pair_cg2); const int* gid = E.I(DI_cg_geomid);
      const int ncon = E.p_misc()[MISC_NCON];
      int first = -1;
      for (int c = 0; c < ncon && first < 0; c++) {
        int p = __float_as_int(E.x_con()[B2_CON_STRIDE * c + 13]);
        int k1 = tp.aux_i[gid[pc1[p]]], k2 = tp.aux_i[gid[pc2[p]]];
        if (k1 == 100) { if (k2 >= 0 && k2 < 9) first = k2; }
        else if (k2 == 100) { if (k1 >= 0 && k1 < 9) first = k1; }
      }
      int held = ti[3];
      if (first >= 0) {
        if (held < 0) { ti[3] = first; ti[4] = 1; ti[5 + first] = 1; }
        else ti[4] = 2;
      } else if (held >= 0) {
        const float* xp = E.p_xpos() + 3 * tp.ids[held]; const float* tg = tp.aux_f + 3 * held;
        float dx = xp[0] - tg[0], dy = xp[1] - tg[1], dz = xp[2] - tg[2];
        if (sqrtf(dx * dx + dy * dy + dz * dz) < 0.002f) { ti[1] |= 1 << held; ti[5 + held] = 2; ti[4] = 4; }
        else { ti[5 + held] = 3; ti[4] = 0; }
        ti[3] = -1;
      } else ti[4] = 0;
    }
    E.sync();
  }
  template <class EN> __device__ static void observe(EN& E, const TaskParams& tp, float* obs) {
    const int* ti = E.p_ti();
    const float mf = max_contact_force(E);
    const int eb = tp.ids[9]; const float* xm = E.p_xmat() + 9 * eb; const float* so = tp.aux_f + 27;
    int nprog = __popc((unsigned)ti[1] & 0x1ffu);
    for (int i = E.lane; i < OBS; i += 32) {
      float v = 0.f;
      if (i < 7) v = E.p_qpos()[i];
      else if (i < 14) v = E.p_qvel()[i - 7];
      else if (i == 14) v = (E.p_qpos()[7] + E.p_qpos()[8]) / 2.0f * 1000.0f;
      else if (i == 15) v = mf;
      else if (i < 19) { int k = i - 16; v = E.p_xpos()[3 * eb + k] + xm[3 * k] * so[0] + xm[3 * k + 1] * so[1] + xm[3 * k + 2] * so[2]; }
      else if (i < 23) v = i == 19 ? 1.f : 0.f;
      else if (i < 79) { int k = (i - 23) / 7, f = (i - 23) % 7; v = f < 3 ? E.p_xpos()[3 * tp.ids[k] + f] : (f == 3 ? 1.f : 0.f); }   // components 0..7 (component 8 is overwritten below)
      else if (i < 87) v = (float)((ti[1] >> (i - 79)) & 1);
      else if (i == 87) v = ti[3] >= 0 ? 1.f : 0.f;
      else if (i == 88) v = (float)ti[3];
      else if (i < 104) v = 0.5f;
      else if (i == 108) v = (float)nprog / 9.0f * 100.0f;
      else if (i == 109) v = (float)ti[4];
      obs[i] = v;
    }
  }
  template <class EN> __device__ static float reward_and_done(EN& E, const TaskParams& tp, const float*, int* ti, float* tf,
                                          int* terminated, int* truncated) {
    float reward = -10.0f;
    const int held = ti[3];
    if (ti[4] == 1 && held >= 0) reward += 1000.f;
    int ndropped = 0;
    for (int k = 0; k < 9; k++) {
      if (((ti[1] >> k) & 1) && ti[5 + k] == 2) reward += (k == 0 || k == 5) ? 2000.f : (k >= 1 && k <= 4 ? 500.f : 1000.f);
      if (ti[5 + k] == 3) ndropped++;
    }
    if (held >= 0) {
      const float* xp = E.p_xpos() + 3 * tp.ids[held]; const float* tg = tp.aux_f + 3 * held;
      float dx = xp[0] - tg[0], dy = xp[1] - tg[1], dz = xp[2] - tg[2], dist = sqrtf(dx * dx + dy * dy + dz * dz);
      if (dist < 0.05f) reward += 300.f * (1.f - dist / 0.05f);
    }
    const float mf = max_contact_force(E);
    if (mf > 50.0f) reward -= 5000.f; else if (mf < 10.0f) reward += 200.f;
    float s = 0.f;
    for (int i = 0; i < 7; i++) s += fabsf(E.p_qvel()[i]);
    reward += -s * 10.f;
    reward -= 2000.f * (float)ndropped;
    const bool all = (ti[1] & 0x1ff) == 0x1ff;
    if (all) reward += 10000.f;
    const float lo[7] = {-3.14f, -2.36f, -2.97f, -3.14f, -2.09f, -3.14f, -3.14f}, hi[7] = {3.14f, 0.78f, 2.97f, 3.14f, 2.09f, 3.14f, 3.14f};
    int term = all ? 1 : 0;
    for (int i = 0; i < 7; i++) { float q = E.p_qpos()[i]; if (q < lo[i] * 0.95f || q > hi[i] * 0.95f) term = 1; }
    *terminated = term; *truncated = ti[0] >= MAX_STEPS;
    tf[0] += reward;
    return reward;
  }
};

}  // namespace b2
