// b2_tasks.cuh -- per-task clip/ctrl, observation, reward, termination and auto-reset fused on the end of the step.
// Each task restates one reference env's order of operations including its index-aliasing
// quirks; all reads of xpos/contacts use the kinematics of the *last forward pass* (SURVEY F9).
#pragma once
#include "b2_engine.cuh"

namespace b2 {

enum { TASK_NONE = 0, TASK_QUADRUPED_PARKOUR = 1 };

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
};

// ------------------------------------------------------------------------------------------------ quadruped
// quadruped_parkour_env/parkour_env.py: step :356-394, _get_observation :396-468, _calculate_reward :646-725,
// _is_terminated :727-755, reset :314-354, _randomize_obstacles :757-774, _update_dynamic_obstacles :776-795.
// ti: [0] step_count [1] checkpoint bitset (6 checkpoints + 12 obstacles) [2] fall_count [3] stuck_counter [4] episode id
// tf: [0] episode_reward [1] last_position.x [2] max_forward_progress
// ids: [0] torso body, [1..4] foot bodies, [5] platform_slide joint id, [6] pendulum_swing joint id,
//      [7] platform_motor actuator id, [8] pendulum_motor actuator id
struct QuadrupedTask {
  static constexpr int OBS = 95, ACT = 16, FRAME_SKIP = 10, SETTLE = 10, MAX_STEPS = 6000, NTI = 8, NTF = 4;

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
        const int* pc1 = E.I(DI_pair_cg1); const int* pc2 = E.I(DI_pair_cg2); const int* gid = E.I(DI_cg_geomid);
        for (int c = 0; c < ncon; c++) {
          int p = __float_as_int(E.p_con()[B2_CON_STRIDE * c + 13]);
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
      const int* pc1 = E.I(DI_pair_cg1); const int* pc2 = E.I(DI_pair_cg2); const int* gid = E.I(DI_cg_geomid);
      for (int f = 0; f < 4; f++) {
        int fid = tp.ids[1 + f];
        for (int c = 0; c < ncon; c++) {
          int p = __float_as_int(E.p_con()[B2_CON_STRIDE * c + 13]);
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

  template <class EN> __device__ static void post_physics(EN& E, const TaskParams& tp, const int* ti) {
    // _update_dynamic_obstacles: uses the pre-increment step counter; takes effect on the next step
    if (E.lane == 0) {
      float t = (float)ti[0] * 0.01f;
      E.p_ctrl()[tp.ids[7]] = 50.0f * sinf(0.5f * t);
      E.p_ctrl()[tp.ids[8]] = 100.0f * sinf(0.3f * t);
    }
    E.sync();
  }
};

}  // namespace b2
