// b2_api.cu -- the step kernel and the C-ABI declared in include/b2env.h (libb2env.so).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>
#include <string>
#include <vector>
#include "../../include/b2env.h"
#include "b2_engine.cuh"
#include "b2_tasks.cuh"

using namespace b2;

static thread_local std::string g_err;
static std::atomic<unsigned long long> g_launches{0};
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(B2_ERR_CUDA, std::string(#x ": ") + cudaGetErrorString(e_)); } while (0)

struct B2Model {
  int device; DevModel dm; int* d_ints; float* d_flts;
  std::vector<int> h_ints;
  bool condim6 = false;     // some contact pair has condim 6 (needs a kernel built with 10-row pyramids)
};
struct B2Batch {
  B2Model* m; int n_envs, W; TaskParams tp; BatchView v; size_t smem;
  DevModel dm;        // the model as this batch's kernels see it (pair tables staged or cold)
  double* epstat;     // [N][4] episodes, return_sum, length_sum, (spare)
  double* d_stats;
  // b2_step_host staging: ONE packed result block [obs | rew | term | trunc] on each side, so a step is one H2D copy, one
  // launch and one D2H copy; the pinned host side is also handed out (b2_host_buffers) for zero-copy callers
  float *h_act, *d_act; char *h_out, *d_out; size_t out_bytes, off_rew, off_term, off_trunc;
  float *h_obs, *h_rew; uint8_t *h_term, *h_trunc;           // views into h_out
  float *d_obs, *d_rew, *d_inject; uint8_t *d_term, *d_trunc; // views into d_out (d_inject separate)
  cudaStream_t own_stream; cudaEvent_t ev_order; cudaStream_t last_stream; bool pending;   // caller-stream work not yet ordered before own_stream
  int episode_slot;   // index of the episode id in ti (RNG stream key), -1 none
  // b2_rollout: the T launches of a rollout as one instantiated CUDA graph, cached while T and the buffers stay the same
  cudaGraphExec_t roll_exec; int roll_T; const void* roll_key[6]; unsigned long long roll_launches;
  int obs_dim, act_dim, nti, ntf, ninj;
  int num_sms;
  int* d_order; bool lpt;   // longest-last-step-first queue order, rebuilt before every b2_step
};

// ------------------------------------------------------------------------------------------------ kernel
// Persistent CTAs: grid = min(#SMs, ceil(n_envs / E)) CTAs of E teams of W warps, one CTA per SM (shared-memory bound).  The
// model tables are staged once per CTA; every team then pulls env indices from a global work queue (one atomicAdd per env)
// until it is empty, so a team whose env is slow -- auto-reset with its settle steps, a wide-tier pass, a NaN retry -- holds
// up nobody: its CTA-mates keep pulling work, and there is no wave tail (r01: 683 CTAs = 4.61 waves, retirement by CTA).
// The last team to leave re-arms the queue for the next launch, so a launch is one kernel node (CUDA-graph friendly).
// Shared memory: [model | E workspaces | mbarrier].  The task hooks are warp-level code run by warp 0 of the team.
template <class Task, int W>
__device__ __forceinline__ void b2_env_body(const DevModel& P, const BatchView& B, const TaskParams& tp, int mode, double* epstat,
                                            const float* inject, int team, int env);

template <class Task, int W>
__global__ void __launch_bounds__(32 * W * Task::MAX_EPB, 1)
b2_env_kernel(const __grid_constant__ DevModel P, const __grid_constant__ BatchView B, const __grid_constant__ TaskParams tp,
              int mode, double* epstat, const float* inject) {
  constexpr int TEAM = 32 * W;
  const int team = threadIdx.x / TEAM, tl = threadIdx.x % TEAM;
  uint64_t* bar = (uint64_t*)(b2_smem + B.model_floats + B.envs_per_block * B.ws_floats);
  volatile int* active = (volatile int*)(bar + 1);       // teams of this CTA that may still pull work (lockstep mode)
  if (threadIdx.x == 0) *active = B.envs_per_block;
  stage_model(P, (int*)b2_smem, b2_smem + r4(P.n_ints_staged), bar);
  volatile int* slot = (volatile int*)(b2_smem + B.model_floats + team * B.ws_floats + B.off.misc) + MISC_ENV;
  for (;;) {
    if (tl == 0) *slot = atomicAdd(B.queue, 1);
    asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(TEAM) : "memory");
    const int idx = *slot;
    if (idx >= B.n_envs) break;
    const int env = B.order ? B.order[idx] : B.first_env + idx;
    if (mode == MODE_RESET && B.reset_mask && !B.reset_mask[env]) continue;
    const long long t0 = clock64();
    b2_env_body<Task, W>(P, B, tp, mode, epstat, inject, team, env);
    if (tl == 0 && mode == MODE_STEP) B.cost[env] = (unsigned)((clock64() - t0) >> 8);
  }
  if (B.lockstep) {
    // a team that is out of work keeps arriving at the CTA barrier of the forward passes until every team of the CTA is done
    if (tl == 0) atomicSub((int*)active, 1);
    for (;;) {
      asm volatile("bar.sync 0;" ::: "memory");
      if (*active == 0) break;
    }
  }
  if (tl == 0) {      // last team out re-arms the queue
    __threadfence();
    const int total = (int)gridDim.x * B.envs_per_block;
    if (atomicAdd(B.queue + 1, 1) == total - 1) { B.queue[1] = 0; __threadfence(); B.queue[0] = 0; }
  }
}

template <class Task, int W>
__device__ __forceinline__ void b2_env_body(const DevModel& P, const BatchView& B, const TaskParams& tp, int mode, double* epstat,
                                            const float* inject, int team, int env) {
  Engine<W, Task::PGS_HOIST, Task::COOP_MIN, Task::COLD_PAIRS, Task::DYN_ISLANDS, Task::SOLVER, Task::CONDIM6, Task::NEWTON_TEAM_ND, Task::CONVEX_PAIRS> E(P, B, B.model_floats + team * B.ws_floats, team, env);
  const int lane = E.lane, tl = E.tl; const bool w0 = E.wl == 0;
  constexpr int TEAM = 32 * W;
  E.team_sync();      // the previous env's stores to this workspace are done
  const int nq = P.dim[DD_nq], nv = P.dim[DD_nv], nu = P.dim[DD_nu];
  unsigned long long* ctr = B.counters + (size_t)env * CTR_COUNT;
  int* s_ti = E.p_ti(); float* s_tf = E.p_tf(); float* s_act = E.p_act();
  // ---- load the env's state rows: 128-bit accesses, consecutive threads on consecutive 16-byte words of one row (rows are padded
  // to 16 bytes in HBM and in the workspace, so the padding words simply travel along)
  {
    auto ld4 = [&](float* dst, const float* src, int nwords) {
      const float4* s4 = reinterpret_cast<const float4*>(src); float4* d4 = reinterpret_cast<float4*>(dst);
      for (int i = tl; i < (nwords >> 2); i += TEAM) d4[i] = s4[i];
    };
    ld4(E.p_qpos(), B.qpos + (size_t)env * B.nqp, B.nqp);
    ld4(E.p_qvel(), B.qvel + (size_t)env * B.nvp, B.nvp); ld4(E.p_warm(), B.warm + (size_t)env * B.nvp, B.nvp);
    ld4(E.p_qapp(), B.qfrc_applied + (size_t)env * B.nvp, B.nvp);
    if (nu > 0) ld4(E.p_ctrl(), B.ctrl + (size_t)env * B.nup, B.nup);
    if (tl < 8) E.p_xfrc()[tl] = 0.f;
    if (tl == 0) { *E.p_time() = B.time[env]; E.p_misc()[MISC_NCON] = 0; E.p_misc()[MISC_NEFC] = 0; E.p_misc()[MISC_FLAG] = 0; E.p_misc()[MISC_DONE] = 0; E.p_misc()[MISC_WIDE] = 0; }
    if (!Task::DYN_ISLANDS && w0) E.static_islands();
    if (Task::NTI % 4 == 0 && Task::NTF % 4 == 0) {
      if (Task::NTI) ld4(reinterpret_cast<float*>(s_ti), reinterpret_cast<const float*>(B.ti + (size_t)env * B.nti), Task::NTI);
      if (Task::NTF) ld4(s_tf, B.tf + (size_t)env * B.ntf, Task::NTF);
    } else {
      for (int i = tl; i < Task::NTI; i += TEAM) s_ti[i] = B.ti[(size_t)env * B.nti + i];
      for (int i = tl; i < Task::NTF; i += TEAM) s_tf[i] = B.tf[(size_t)env * B.ntf + i];
    }
  }
  E.team_sync();

  // one loop, one call site of the physics: [reset ->] n sub-steps -> task epilogue [-> auto-reset -> settle steps]
  // (nfwd: a reset that ends with mj_forward instead of settle steps, Task::RESET_FORWARD)
  int nsub = 0, nfwd = 0, stage = 0;
  if (mode == MODE_PHYS) nsub = B.nsub;
  else if (mode == MODE_FORWARD) nsub = 1;
  else if (mode == MODE_RESET) {
    if (w0) Task::reset_state(E, tp, B, env, s_ti, s_tf, inject ? inject + (size_t)Task::NINJ * env : nullptr);
    E.team_sync(); nsub = Task::SETTLE; nfwd = Task::RESET_FORWARD ? 1 : 0; stage = 1;
  } else {
    if (w0) { Task::apply_action(E, tp, B.action + (size_t)env * B.act_dim, s_act); Task::pre_physics(E, tp, s_ti, s_tf); }
    E.team_sync(); nsub = Task::FRAME_SKIP;
  }
  while (true) {
    if (nsub > 0 || (Task::RESET_FORWARD && nfwd > 0)) { E.step(ctr, nsub > 0 && mode != MODE_FORWARD); if (nsub > 0) nsub--; else nfwd--; continue; }
    if (mode == MODE_PHYS || mode == MODE_FORWARD) break;
    if (stage == 0) {   // end of the control step
      if (w0) {
        Task::post_physics(E, tp, s_ti, s_tf);
        Task::observe(E, tp, B.obs + (size_t)env * B.obs_dim);
        E.sync();
        if (lane == 0) {
          int term = 0, trunc = 0;
          float rew = Task::reward_and_done(E, tp, s_act, s_ti, s_tf, &term, &trunc);
          E.p_misc()[MISC_DONE] = term | (trunc << 1);
          B.reward[env] = rew; B.term[env] = (uint8_t)term; B.trunc[env] = (uint8_t)trunc;
        }
      }
      E.team_sync();
      if (!E.p_misc()[MISC_DONE]) break;
      // same-step auto-reset: keep the terminal observation, account the episode, start the next one
      if (w0) {
        if (B.final_obs) for (int i = lane; i < Task::OBS; i += 32) B.final_obs[(size_t)env * B.obs_dim + i] = B.obs[(size_t)env * B.obs_dim + i];
        if (lane == 0) {
          epstat[4 * env + 0] += 1.0; epstat[4 * env + 1] += (double)s_tf[0]; epstat[4 * env + 2] += (double)s_ti[0];
          atomicAdd(&ctr[CTR_EPISODES], 1ull);
        }
        // the finished episode's task state and body positions, for the terminal step's info (the reference reports the
        // totals of the episode that just ended, e.g. quadruped_parkour_env/parkour_env.py:797-813)
        if (B.final_ti) {
          for (int i = lane; i < Task::NTI; i += 32) B.final_ti[(size_t)env * B.nti + i] = s_ti[i];
          for (int i = lane; i < Task::NTF; i += 32) B.final_tf[(size_t)env * B.ntf + i] = s_tf[i];
          const int n3 = 3 * P.dim[DD_nbody];
          for (int i = lane; i < n3; i += 32) B.final_xpos[(size_t)env * n3 + i] = E.p_xpos()[i];
        }
        E.sync();
        Task::reset_state(E, tp, B, env, s_ti, s_tf, nullptr);
      }
      E.team_sync();
      nsub = Task::SETTLE; nfwd = Task::RESET_FORWARD ? 1 : 0; stage = 1;
      if (nsub == 0 && nfwd == 0) { if (w0) { Task::after_settle(E, tp, s_ti, s_tf); Task::observe(E, tp, B.obs + (size_t)env * B.obs_dim); } break; }
    } else {            // end of the settle steps of a reset
      if (w0) { Task::after_settle(E, tp, s_ti, s_tf); Task::observe(E, tp, B.obs + (size_t)env * B.obs_dim); }
      break;
    }
  }
  E.team_sync();
  // ---- optional exports of the last forward pass
  if (B.c_ncon) {
    int ncon = E.p_misc()[MISC_NCON];
    const int* pc1 = E.PI(DI_pair_cg1); const int* pc2 = E.PI(DI_pair_cg2); const int* gid = E.I(DI_cg_geomid);
    if (tl == 0) B.c_ncon[env] = ncon;
    for (int c = tl; c < ncon && c < B.c_cap; c += TEAM) {
      int p = __float_as_int(E.x_con()[B2_CON_STRIDE * c + 13]);
      B.c_geom[((size_t)env * B.c_cap + c) * 2] = gid[pc1[p]]; B.c_geom[((size_t)env * B.c_cap + c) * 2 + 1] = gid[pc2[p]];
      B.c_dist[(size_t)env * B.c_cap + c] = E.x_con()[B2_CON_STRIDE * c];
    }
  }
  if (B.xpos_out) { int n = 3 * P.dim[DD_nbody]; for (int i = tl; i < n; i += TEAM) B.xpos_out[(size_t)env * n + i] = E.p_xpos()[i]; }
  if (B.debug_out) {
    // layout: qfs | qas | qfc | qacc (nv each) | M (nM) | ncon nefc iters 0 | row_f | row_b | row_R | row_res (row_cap each)
    float* o = B.debug_out + (size_t)env * B.debug_n; int nM = P.dim[DD_nM], k = 0;
    auto put = [&](const float* src, int n) { for (int i = tl; i < n; i += TEAM) if (k + i < B.debug_n) o[k + i] = src[i]; k += n; };
    put(E.p_qfs(), nv); put(E.p_qas(), nv); put(E.p_qfc(), nv); put(E.p_qacc(), nv); put(E.p_M(), nM);
    if (tl == 0 && k + 4 <= B.debug_n) { o[k] = (float)E.p_misc()[MISC_NCON]; o[k + 1] = (float)E.p_misc()[MISC_NEFC]; o[k + 2] = (float)E.p_misc()[MISC_ITERS]; o[k + 3] = (float)E.p_misc()[MISC_WIDE]; }
    k += 4;
    put(E.x_row_f(), B.row_cap); put(E.x_row_b(), B.row_cap); put(E.x_row_R(), B.row_cap); put(E.x_row_res(), B.row_cap);
  }
  // ---- store state (128-bit, like the load)
  if (mode == MODE_FORWARD && B.forward_saves_warm && !B.warm_once) {
    float* gw = B.warm + (size_t)env * B.nvp;
    for (int i = tl; i < nv; i += TEAM) gw[i] = E.p_warm()[i];
  }
  if (mode != MODE_FORWARD) {
    auto st4 = [&](float* dst, const float* src, int nwords) {
      const float4* s4 = reinterpret_cast<const float4*>(src); float4* d4 = reinterpret_cast<float4*>(dst);
      for (int i = tl; i < (nwords >> 2); i += TEAM) d4[i] = s4[i];
    };
    st4(B.qpos + (size_t)env * B.nqp, E.p_qpos(), B.nqp);
    st4(B.qvel + (size_t)env * B.nvp, E.p_qvel(), B.nvp); st4(B.warm + (size_t)env * B.nvp, E.p_warm(), B.nvp);
    st4(B.qfrc_applied + (size_t)env * B.nvp, E.p_qapp(), B.nvp);
    if (nu > 0) st4(B.ctrl + (size_t)env * B.nup, E.p_ctrl(), B.nup);
    if (tl == 0) B.time[env] = *E.p_time();
    if (Task::NTI % 4 == 0 && Task::NTF % 4 == 0) {
      if (Task::NTI) st4(reinterpret_cast<float*>(B.ti + (size_t)env * B.nti), reinterpret_cast<const float*>(s_ti), Task::NTI);
      if (Task::NTF) st4(B.tf + (size_t)env * B.ntf, s_tf, Task::NTF);
    } else {
      for (int i = tl; i < Task::NTI; i += TEAM) B.ti[(size_t)env * B.nti + i] = s_ti[i];
      for (int i = tl; i < Task::NTF; i += TEAM) B.tf[(size_t)env * B.ntf + i] = s_tf[i];
    }
  }
}

// a task that only runs physics (b2_physics_step / b2_forward on a model without task logic)
struct NoTask {
  static constexpr int OBS = 0, ACT = 0, FRAME_SKIP = 1, SETTLE = 0, MAX_STEPS = 0, NTI = 0, NTF = 0, NINJ = 1, KEEP_FRAMES = 0, XFRC_SLOT = -1, COOP_MIN = 32, ARENA_ROWS = 80, CON_CAP = 32, ARENA_SPAN = 0, MAX_EPB = 6, EPISODE_SLOT = -1, LOCKSTEP = 0, NEWTON_TEAM_ND = 16, ARENA_FLOATS = 0;
  static constexpr int SOLVER = -1;
  static constexpr bool CONDIM6 = true, RESET_FORWARD = false, PGS_HOIST = true, COLD_PAIRS = false, DYN_ISLANDS = true, CONVEX_PAIRS = true;
  template <class EN> __device__ static void apply_action(EN&, const TaskParams&, const float*, float*) {}
  template <class EN> __device__ static void pre_physics(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void after_settle(EN&, const TaskParams&, int*, float*) {}
  template <class EN> __device__ static void reset_state(EN& E, const TaskParams&, const BatchView&, int, int*, float*, const float*) {
    E.reset_data(); if (E.lane == 0) *E.p_time() = 0.f; E.sync();
  }
  template <class EN> __device__ static void observe(EN&, const TaskParams&, float*) {}
  template <class EN> __device__ static float reward_and_done(EN&, const TaskParams&, const float*, int*, float*, int* a, int* b) { *a = 0; *b = 0; return 0.f; }
  template <class EN> __device__ static void post_physics(EN&, const TaskParams&, int*, float*) {}
};

// Longest-processing-time-first order of the work queue: envs whose last control step was expensive (a fallen robot with 40
// contacts sweeps three times the rows of a standing one, a blown-up one runs in the wide tier) tend to be expensive again,
// and a long job pulled last is a tail every other SM waits for.  One CTA buckets the envs by last-step cost relative to the
// mean (8 buckets, most expensive first); order inside a bucket is by env index.  Runs before every b2_step launch (~60 us cold under ncu, 0.4 % of a step).
__global__ void b2_order_kernel(const unsigned* cost, int n, int* order) {
  __shared__ unsigned long long ssum; __shared__ int cnt[8], base[8];
  if (threadIdx.x == 0) ssum = 0ull;
  if (threadIdx.x < 8) cnt[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long loc = 0;
  for (int e = threadIdx.x; e < n; e += blockDim.x) loc += cost[e];
  atomicAdd(&ssum, loc);
  __syncthreads();
  const float mean = (float)ssum / (float)n + 1.0f;
  auto bucket = [mean](unsigned c) { const float r = (float)c / mean; return r > 3.0f ? 0 : r > 2.2f ? 1 : r > 1.7f ? 2 : r > 1.35f ? 3 : r > 1.1f ? 4 : r > 0.9f ? 5 : r > 0.7f ? 6 : 7; };
  // stable within a bucket: each thread owns a contiguous slice of envs
  const int per = (n + blockDim.x - 1) / blockDim.x, e0 = threadIdx.x * per, e1 = min(n, e0 + per);
  int mine[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int e = e0; e < e1; e++) mine[bucket(cost[e])]++;
  __shared__ int slice[8][1024];
  for (int k = 0; k < 8; k++) slice[k][threadIdx.x] = mine[k];
  __syncthreads();
  if (threadIdx.x < 8) { int acc = 0; for (int t = 0; t < (int)blockDim.x; t++) { int v = slice[threadIdx.x][t]; slice[threadIdx.x][t] = acc; acc += v; } cnt[threadIdx.x] = acc; }
  __syncthreads();
  if (threadIdx.x == 0) { int acc = 0; for (int k = 0; k < 8; k++) { base[k] = acc; acc += cnt[k]; } }
  __syncthreads();
  int pos[8];
  for (int k = 0; k < 8; k++) pos[k] = base[k] + slice[k][threadIdx.x];
  for (int e = e0; e < e1; e++) { const int k = bucket(cost[e]); order[pos[k]++] = e; }
}

__global__ void b2_stats_kernel(const unsigned long long* counters, const double* epstat, int n, double* out) {
  __shared__ double acc[16];
  if (threadIdx.x < 16) acc[threadIdx.x] = 0.0;
  __syncthreads();
  double loc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    loc[0] += epstat[4 * e]; loc[1] += epstat[4 * e + 1]; loc[2] += epstat[4 * e + 2];
    const unsigned long long* c = counters + (size_t)e * CTR_COUNT;
    loc[3] += (double)c[CTR_NAN_RESET]; loc[4] += (double)c[CTR_CON_DROPPED]; loc[5] += (double)c[CTR_ROW_DROPPED];
    loc[6] += (double)c[CTR_ARENA_OVERFLOW]; loc[7] += (double)c[CTR_SOLVER_ITERS]; loc[8] += (double)c[CTR_SUBSTEPS]; loc[9] += (double)c[CTR_ARENA_SPILL]; loc[10] += (double)c[CTR_WIDE]; loc[11] += (double)c[CTR_WIDE_ROWS];
  }
  for (int k = 0; k < 12; k++) atomicAdd(&acc[k], loc[k]);
  __syncthreads();
  if (threadIdx.x < 16) atomicAdd(&out[threadIdx.x], acc[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ host side
// The dynamic shared-memory opt-in is a per-device function attribute: it is set when a batch is created (on the batch's
// device), never on the launch path, so that b2_step can be captured into a CUDA graph.
template <class Task>
static int configure_task(B2Batch* b) {
  // the attribute belongs to the function (per device), not to the batch: always opt in to the full 227 KB, so that a later,
  // smaller batch of the same task cannot lower the limit under an earlier one
  (void)b;
  CK(cudaFuncSetAttribute(b2_env_kernel<Task, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return B2_OK;
}
template <class Task>
static int launch_task(B2Batch* b, int mode, const float* inject, cudaStream_t s) {
  int E = b->v.envs_per_block, grid = (b->v.n_envs + E - 1) / E;
  if (grid > b->num_sms) grid = b->num_sms;            // persistent CTAs, one per SM; teams pull envs from the work queue
  b2_env_kernel<Task, 3><<<grid, 32 * 3 * E, b->smem, s>>>(b->dm, b->v, b->tp, mode, b->epstat, inject);
  g_launches++;
  CK(cudaGetLastError());
  return B2_OK;
}
#ifdef B2_ONLY_TASK      /* bring-up builds (tools/phase_timing.py, A/B experiments): one task's kernel only, compiles in a fraction of the time */
#define B2_FOR_TASK(b, CALL) return CALL(B2_ONLY_TASK)
#else
#define B2_FOR_TASK(b, CALL) \
  switch ((b)->tp.task) { \
    case TASK_NONE: return CALL(NoTask); \
    case TASK_QUADRUPED_PARKOUR: return CALL(QuadrupedTask); \
    case TASK_HUMANOID_DANCING: return CALL(DancingTask); \
    case TASK_HUMANOID_SOCCER: return CALL(SoccerTask); \
    case TASK_BIPEDAL_RESCUE: return CALL(RescueTask); \
    case TASK_HUMANOID_CONSTRUCTION: return CALL(ConstructionTask); \
    case TASK_HUMANOID_MARTIAL_ARTS: return CALL(MartialArtsTask); \
    case TASK_ROBOTIC_ARM_ASSEMBLY: return CALL(ArmTask); \
  } \
  return fail(B2_ERR_UNSUPPORTED, "unknown task id")
#endif
static int configure(B2Batch* b) {
#define B2_CALL(T) configure_task<T>(b)
  B2_FOR_TASK(b, B2_CALL);
#undef B2_CALL
}
// Work issued on a caller stream and work issued on own_stream (b2_step_host) touch the same state rows: whenever the stream
// changes, the new one first waits for an event recorded on the previous one.
static int order_after_last(B2Batch* b, cudaStream_t s) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) return B2_OK;    // inside a graph capture: the capturer orders
  if (b->pending && b->last_stream != s) { CK(cudaEventRecord(b->ev_order, b->last_stream)); CK(cudaStreamWaitEvent(s, b->ev_order, 0)); }
  b->last_stream = s; b->pending = true;
  return B2_OK;
}
static int launch(B2Batch* b, int mode, const float* inject, cudaStream_t s) {
  CK(cudaSetDevice(b->m->device));
  { int rc = order_after_last(b, s); if (rc) return rc; }
  if (b->v.n_envs <= 0 || b->v.first_env + b->v.n_envs > b->n_envs) { b->v.first_env = 0; b->v.n_envs = b->n_envs; }
  b->v.order = nullptr;
  if (mode == MODE_STEP && b->lpt && b->v.first_env == 0 && b->v.n_envs == b->n_envs && b->n_envs >= 4 * b->num_sms) {
    b2_order_kernel<<<1, 1024, 0, s>>>(b->v.cost, b->n_envs, b->d_order);
    g_launches++;
    b->v.order = b->d_order;
  }
#define B2_CALL(T) launch_task<T>(b, mode, inject, s)
  B2_FOR_TASK(b, B2_CALL);
#undef B2_CALL
}

extern "C" {

void b2_batch_destroy(B2Batch* b);
const char* b2_last_error(void) { return g_err.c_str(); }
unsigned long long b2_launch_count(void) { return g_launches.load(); }
/* bring-up hook (not in b2env.h): per-phase clock64 sums when the library is built with -DB2_PHASE_TIMING */
int b2_phase_cycles(B2Batch* b, unsigned long long* out16) {
  if (!b || !out16) return B2_ERR_ARG;
  cudaSetDevice(b->m->device); cudaDeviceSynchronize();
  return cudaMemcpy(out16, b->v.phase_cycles, 32 * 8, cudaMemcpyDeviceToHost) == cudaSuccess ? B2_OK : B2_ERR_CUDA;   /* 32 entries */
}

int b2_model_create(const int32_t* ints, int n_ints, const double* flts, int n_flts, int device, B2Model** out) {
  if (!ints || !flts || !out) return fail(B2_ERR_ARG, "null argument");
  if (n_ints < 4 || ints[0] != (int)B2DEV_MAGIC || ints[1] != B2DEV_N_INT_FIELDS || ints[2] != B2DEV_N_FLT_FIELDS)
    return fail(B2_ERR_LAYOUT, "packed device model does not match include/b2_device_layout.h");
  if ((n_ints & 3) || (n_flts & 3)) return fail(B2_ERR_LAYOUT, "buffers must be padded to 4 elements");
  int ndev = 0; CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(B2_ERR_ARG, "no such CUDA device");
  CK(cudaSetDevice(device));
  B2Model* m = new B2Model(); m->device = device; m->h_ints.assign(ints, ints + n_ints);
  DevModel& dm = m->dm; memset(&dm, 0, sizeof(dm));
  for (int k = 0; k < B2DEV_N_INT_FIELDS; k++) dm.ioff[k] = ints[4 + 2 * k];
  for (int k = 0; k < B2DEV_N_FLT_FIELDS; k++) dm.foff[k] = ints[4 + 2 * (B2DEV_N_INT_FIELDS + k)];
  const int* dims = ints + dm.ioff[DI_dims];
  int ndims = ints[5 + 2 * DI_dims];
  if (ndims != DD_COUNT) { delete m; return fail(B2_ERR_LAYOUT, "dims field has the wrong length"); }
  for (int k = 0; k < DD_COUNT; k++) dm.dim[k] = dims[k];
  const double* opt = flts + dm.foff[DF_opt];
  for (int k = 0; k < DO_COUNT; k++) dm.opt[k] = (float)opt[k];
  if (dm.dim[DD_integrator] != 0 && dm.dim[DD_integrator] != 1) { delete m; return fail(B2_ERR_UNSUPPORTED, "integrator must be Euler or RK4"); }
  if (dm.dim[DD_solver] != 0 && dm.dim[DD_solver] != 2) { delete m; return fail(B2_ERR_UNSUPPORTED, "solver must be PGS or Newton"); }
  if (dm.dim[DD_ntree] > B2_MAX_ISLANDS) { delete m; return fail(B2_ERR_UNSUPPORTED, "more than 16 kinematic trees"); }
  for (int r = 0; r < dm.dim[DD_nprm]; r++) {
    double cd = flts[dm.foff[DF_prm] + B2DEV_PRM_STRIDE * r + 14];
    if (cd == 6.0) m->condim6 = true;
    else if (cd != 3.0 && dm.dim[DD_npair] > 0) { delete m; return fail(B2_ERR_UNSUPPORTED, "contact pairs must be condim 3 or 6"); }
  }
  std::vector<float> f32(n_flts);
  for (int i = 0; i < n_flts; i++) f32[i] = (float)flts[i];
  CK(cudaMalloc(&m->d_ints, sizeof(int) * n_ints)); CK(cudaMalloc(&m->d_flts, sizeof(float) * n_flts));
  CK(cudaMemcpy(m->d_ints, ints, sizeof(int) * n_ints, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(m->d_flts, f32.data(), sizeof(float) * n_flts, cudaMemcpyHostToDevice));
  dm.ints = m->d_ints; dm.flts = m->d_flts; dm.n_ints = n_ints; dm.n_flts = n_flts; dm.n_ints_staged = n_ints;
  *out = m;
  return B2_OK;
}
void b2_model_destroy(B2Model* m) { if (!m) return; cudaSetDevice(m->device); cudaFree(m->d_ints); cudaFree(m->d_flts); delete m; }

int b2_batch_create(B2Model* m, const B2TaskDesc* task, int n_envs, uint64_t seed, int env_offset, const B2BatchOpts* opts,
                    B2Batch** out) {
  if (!m || !out || n_envs <= 0) return fail(B2_ERR_ARG, "bad argument");
  CK(cudaSetDevice(m->device));
  B2Batch* b = new B2Batch(); memset(b, 0, sizeof(*b));
  b->m = m; b->n_envs = n_envs;
  struct Guard { B2Batch* b; ~Guard() { if (b) b2_batch_destroy(b); } } guard{b};      // released on success
  memset(&b->tp, 0, sizeof(b->tp));
  int task_solver = -1, keep_frames = 0; bool task_c6 = true, task_cvx = true; int xfrc_body = -1, arena_rows = 80, task_con_cap = 32, arena_span = 0, max_epb = 6, task_lockstep = 0, task_arena = 0; bool cold = false; b->ninj = 1;
  if (task) { b->tp.task = task->task; memcpy(b->tp.ids, task->ids, sizeof(task->ids)); memcpy(b->tp.act_lo, task->act_lo, sizeof(task->act_lo)); memcpy(b->tp.act_hi, task->act_hi, sizeof(task->act_hi)); memcpy(b->tp.aux_i, task->aux_i, sizeof(task->aux_i)); memcpy(b->tp.aux_f, task->aux_f, sizeof(task->aux_f)); }
  switch (b->tp.task) {
    case TASK_NONE: b->obs_dim = 0; b->act_dim = 0; b->nti = 0; b->ntf = 0; b->episode_slot = -1; break;
#define B2_TASK_DIMS(T) task_solver = T::SOLVER; task_c6 = T::CONDIM6; task_cvx = T::CONVEX_PAIRS; b->obs_dim = T::OBS; b->act_dim = T::ACT; b->nti = T::NTI; b->ntf = T::NTF; b->ninj = T::NINJ; keep_frames = T::KEEP_FRAMES; cold = T::COLD_PAIRS; arena_rows = T::ARENA_ROWS; task_con_cap = T::CON_CAP; arena_span = T::ARENA_SPAN; max_epb = T::MAX_EPB; b->episode_slot = T::EPISODE_SLOT; task_lockstep = T::LOCKSTEP; task_arena = T::ARENA_FLOATS; xfrc_body = T::XFRC_SLOT >= 0 ? b->tp.ids[T::XFRC_SLOT >= 0 ? T::XFRC_SLOT : 0] : -1
    case TASK_QUADRUPED_PARKOUR: B2_TASK_DIMS(QuadrupedTask); break;
    case TASK_HUMANOID_DANCING: B2_TASK_DIMS(DancingTask); break;
    case TASK_HUMANOID_SOCCER: B2_TASK_DIMS(SoccerTask); break;
    case TASK_BIPEDAL_RESCUE: B2_TASK_DIMS(RescueTask); break;
    case TASK_HUMANOID_CONSTRUCTION: B2_TASK_DIMS(ConstructionTask); break;
    case TASK_HUMANOID_MARTIAL_ARTS: B2_TASK_DIMS(MartialArtsTask); break;
    case TASK_ROBOTIC_ARM_ASSEMBLY: B2_TASK_DIMS(ArmTask); break;
    default: delete b; return fail(B2_ERR_UNSUPPORTED, "unknown task id");
  }
  const int* dim = m->dm.dim;
  if (m->condim6 && !task_c6) { return fail(B2_ERR_UNSUPPORTED, "the model has condim-6 pairs but the task kernel is built for 4-row pyramids"); }
  if (task_solver >= 0 && task_solver != dim[DD_solver]) { return fail(B2_ERR_UNSUPPORTED, "the task kernel is compiled for a different <option solver> than the model's"); }
  if (!task_cvx) {      // the task's kernel is built without the convex (MPR) path: the model must not need it
    const int* hi = m->h_ints.data(); const int* c1 = hi + m->dm.ioff[DI_pair_cg1]; const int* c2 = hi + m->dm.ioff[DI_pair_cg2]; const int* ct = hi + m->dm.ioff[DI_cg_type];
    for (int p = 0; p < dim[DD_npair]; p++) {
      const int t1 = ct[c1[p]], t2 = ct[c2[p]];
      if ((t1 == GT_CAPSULE && t2 == GT_CYLINDER) || (t1 == GT_CYLINDER && (t2 == GT_CYLINDER || t2 == GT_BOX)))
        return fail(B2_ERR_UNSUPPORTED, "the model has capsule-cylinder / cylinder-cylinder / cylinder-box candidate pairs but the task kernel is built without the convex path");
    }
  }
  BatchView& v = b->v; memset(&v, 0, sizeof(v));
  v.n_envs = n_envs; v.nqp = r4(dim[DD_nq]); v.nvp = r4(dim[DD_nv]); v.nup = r4(dim[DD_nu] > 0 ? dim[DD_nu] : 1);
  v.nti = b->nti > 0 ? b->nti : 1; v.ntf = b->ntf > 0 ? b->ntf : 1; v.obs_dim = b->obs_dim; v.act_dim = b->act_dim;
  v.warm_once = (opts && opts->warmstart_once_per_step) ? 1 : 0;
  v.seed = seed; v.env_offset = env_offset; v.keep_frames = keep_frames; v.inject_stride = b->ninj; v.xfrc_body = xfrc_body;
  // fixed-capacity buffers (SURVEY App. D suggests 32 contacts for the quadruped); rows: 4 per contact + limits
  int o_epb = opts ? opts->envs_per_block : 0, o_arena = opts ? opts->arena_floats : 0;
  int o_con = opts ? opts->con_cap : 0, o_row = opts ? opts->row_cap : 0;
  b->W = (opts && opts->warps_per_env > 0) ? opts->warps_per_env : 3;
  if (b->W != 3) { return fail(B2_ERR_ARG, "warps_per_env must be 3 (the one-warp-per-env kernels are no longer built)"); }
  v.con_cap = o_con > 0 ? o_con : (dim[DD_maxraw] < task_con_cap ? dim[DD_maxraw] : task_con_cap); if (v.con_cap < 1) v.con_cap = 1;
  v.row_cap = o_row > 0 ? o_row : 4 * v.con_cap + dim[DD_nlim]; if (v.row_cap > B2_ISLAND_ROWS * 4) v.row_cap = B2_ISLAND_ROWS * 4;
  v.row_cap = r4(v.row_cap < 4 ? 4 : v.row_cap);
  int maxspan = 0; const int* inum = m->h_ints.data() + m->dm.ioff[DI_island_dofnum];
  for (int k = 0; k < dim[DD_nisland]; k++) if (inum[k] > maxspan) maxspan = inum[k];
  // narrow-phase capacities: active pairs after the cull, raw contact slots handed out to them
  v.act_cap = r4(dim[DD_npair] < 256 ? dim[DD_npair] : 256); if (v.act_cap < 4) v.act_cap = 4;
  v.raw_cap = dim[DD_maxraw] < 384 ? dim[DD_maxraw] : 384; if (v.raw_cap < 8) v.raw_cap = 8;
  int raw_need = 3 * v.act_cap + 10 * v.raw_cap;
  b->dm = m->dm;
  if (cold) b->dm.n_ints_staged = m->h_ints[3];       // header word 3: where the pair tables start
  int scratch = (dim[DD_solver] == 2 || 32 * dim[DD_nv] <= dead_block_floats(dim, keep_frames)) ? 0 : 32 * dim[DD_nv];   // Newton builds no A
  // default arena: J (rows x widest island) + tiled A for the task's typical row count, plus the A-build scratch
  int typ = arena_rows < B2_ISLAND_ROWS ? arena_rows : B2_ISLAND_ROWS;
  int span = arena_span > 0 && arena_span < maxspan ? arena_span : maxspan;      // dofs of the widest island the arena is sized for
  int arena_default = r4(typ * (span | 1)) + 16 * ((((typ + 3) >> 2) * (((typ + 3) >> 2) + 1)) >> 1) + 64;
  // ARENA_FLOATS: a task whose rows always fit a small arena keeps it small and runs more teams per SM instead (arm: every island
  // is a 6-9 dof body with a few dozen rows; 4 000 floats hold them all, and 4 teams per SM instead of 2 gave +63 %)
  if (o_arena <= 0 && task_arena > 0) { o_arena = task_arena; }
  int arena = (o_arena > 0 ? o_arena : arena_default) + scratch;
  if (arena < raw_need) arena = raw_need;
  v.arena_floats = r4(arena);
  v.ws_floats = ws_layout(dim, v.con_cap, v.row_cap, v.arena_floats, b->nti, b->ntf, &v.off);
  v.model_floats = model_smem_floats(b->dm.n_ints_staged, b->dm.n_flts);
  const int smem_max = 227 * 1024;
  int epb = (smem_max - v.model_floats * 4 - 16) / (v.ws_floats * 4);
  if (epb > max_epb) epb = max_epb;          // the kernel's __launch_bounds__ (register budget) is sized for the task's MAX_EPB
  if (o_epb > 0 && o_epb < epb) epb = o_epb;
  if (epb < 1) { return fail(B2_ERR_UNSUPPORTED, "model needs more than 227 KB of shared memory per env"); }
  v.envs_per_block = epb;
  if (o_arena <= 0) {
    // the CTA's shared memory is allocated anyway: hand what is left over to the arenas (more rows before truncation)
    int spare = ((smem_max - v.model_floats * 4 - 16) / epb - v.ws_floats * 4) / 4;
    spare = (spare / 32) * 32;
    if (spare > 0) { v.arena_floats += spare; v.ws_floats = ws_layout(dim, v.con_cap, v.row_cap, v.arena_floats, b->nti, b->ntf, &v.off); }
  }
  b->smem = ((size_t)v.model_floats + (size_t)epb * v.ws_floats) * 4 + 16;
  size_t N = n_envs;
  // raw contact slots live in the arena during the narrow phase: use what the final arena holds
  if (dim[DD_maxraw] > v.raw_cap) { int fit = (v.arena_floats - 3 * v.act_cap) / 10; v.raw_cap = dim[DD_maxraw] < fit ? dim[DD_maxraw] : fit; if (v.raw_cap > 1024) v.raw_cap = 1024; }
  // wide tier: per-env global workspace for forward passes that exceed the on-chip capacities (nothing is dropped)
  {
    const bool pgs = dim[DD_solver] == 0; const int nvw = (dim[DD_maxspan] + 3) & ~3;
    bool on = !(opts && opts->disable_wide) && dim[DD_npair] > 0 && !(pgs && nvw > 128);
    // the on-chip part of the wide tier must fit the arena: one scratch set + the rings (PGS) / the largest island's H and vectors (Newton)
    int fixed = pgs ? 32 * dim[DD_nv] + 3 * 3 * (8 * nvw + B2_WREC + 4) + 64 : 0;
    if (fixed + 256 > v.arena_floats) on = false;
    if (on) {
      v.w_con_cap = dim[DD_maxraw] < 256 ? dim[DD_maxraw] : 256; if (v.w_con_cap < v.con_cap) v.w_con_cap = v.con_cap;
      int rows = (task_c6 ? 10 : 4) * v.w_con_cap + 2 * dim[DD_nlim]; if (rows > 1024) rows = 1024; if (rows < v.row_cap) rows = v.row_cap;
      v.w_row_cap = r4(rows);
      v.w_arena_floats = pgs ? (v.w_row_cap / 4 + B2_MAX_ISLANDS) * (8 * nvw + B2_WREC) + 16 : r4(v.w_row_cap * ((dim[DD_maxspan] | 1) + 1)) + 16 * B2_MAX_ISLANDS;
      v.wide_stride = wide_layout(v.w_con_cap, v.w_row_cap, v.w_arena_floats, &v.woff);
      CK(cudaMalloc(&v.wide, N * (size_t)v.wide_stride * 4));
    }
  }
  CK(cudaMalloc(&v.qpos, N * v.nqp * 4)); CK(cudaMalloc(&v.qvel, N * v.nvp * 4)); CK(cudaMalloc(&v.warm, N * v.nvp * 4));
  CK(cudaMalloc(&v.qfrc_applied, N * v.nvp * 4)); CK(cudaMalloc(&v.ctrl, N * v.nup * 4)); CK(cudaMalloc(&v.time, N * 4));
  CK(cudaMalloc(&v.ti, N * v.nti * 4)); CK(cudaMalloc(&v.tf, N * v.ntf * 4));
  CK(cudaMalloc(&v.phase_cycles, 32 * 8)); CK(cudaMemset(v.phase_cycles, 0, 32 * 8));
  CK(cudaMalloc(&v.queue, 4 * 4)); CK(cudaMemset(v.queue, 0, 4 * 4));
  CK(cudaMalloc(&v.cost, N * 4)); CK(cudaMemset(v.cost, 0, N * 4)); CK(cudaMalloc(&b->d_order, N * 4));
  b->lpt = !(opts && opts->fifo_queue);
  // per-task default (A/B-measured on B200: +13 % dancing, whose forward passes are instruction-fetch bound and alike in length;
  // -9..-21 % where PGS / Newton iteration counts vary between envs); B2_LOCKSTEP overrides for experiments
  { const char* ls = getenv("B2_LOCKSTEP"); v.lockstep = ls ? atoi(ls) : task_lockstep; }
  CK(cudaDeviceGetAttribute(&b->num_sms, cudaDevAttrMultiProcessorCount, m->device));
  CK(cudaMalloc(&v.counters, N * CTR_COUNT * 8)); CK(cudaMalloc(&b->epstat, N * 4 * 8)); CK(cudaMalloc(&b->d_stats, 16 * 8));
  CK(cudaMemset(v.qvel, 0, N * v.nvp * 4)); CK(cudaMemset(v.warm, 0, N * v.nvp * 4)); CK(cudaMemset(v.qfrc_applied, 0, N * v.nvp * 4));
  CK(cudaMemset(v.ctrl, 0, N * v.nup * 4)); CK(cudaMemset(v.time, 0, N * 4)); CK(cudaMemset(v.ti, 0, N * v.nti * 4));
  CK(cudaMemset(v.tf, 0, N * v.ntf * 4)); CK(cudaMemset(v.counters, 0, N * CTR_COUNT * 8)); CK(cudaMemset(b->epstat, 0, N * 32));
  // qpos <- qpos0
  {
    std::vector<float> q((size_t)N * v.nqp, 0.f);
    std::vector<float> q0(dim[DD_nq]);
    CK(cudaMemcpy(q0.data(), m->d_flts + m->dm.foff[DF_qpos0], sizeof(float) * dim[DD_nq], cudaMemcpyDeviceToHost));
    for (size_t e = 0; e < N; e++) memcpy(&q[e * v.nqp], q0.data(), sizeof(float) * dim[DD_nq]);
    CK(cudaMemcpy(v.qpos, q.data(), q.size() * 4, cudaMemcpyHostToDevice));
  }
  // staging for the host-buffer entry point: one packed result block per side
  size_t od = b->obs_dim > 0 ? b->obs_dim : 1, ad = b->act_dim > 0 ? b->act_dim : 1;
  b->off_rew = (N * od * 4 + 15) & ~(size_t)15; b->off_term = (b->off_rew + N * 4 + 15) & ~(size_t)15;
  b->off_trunc = (b->off_term + N + 15) & ~(size_t)15; b->out_bytes = (b->off_trunc + N + 15) & ~(size_t)15;
  CK(cudaMallocHost(&b->h_act, N * ad * 4)); CK(cudaMallocHost(&b->h_out, b->out_bytes));
  CK(cudaMalloc(&b->d_act, N * ad * 4)); CK(cudaMalloc(&b->d_out, b->out_bytes)); CK(cudaMalloc(&b->d_inject, N * (size_t)b->ninj * 4));
  b->h_obs = (float*)b->h_out; b->h_rew = (float*)(b->h_out + b->off_rew); b->h_term = (uint8_t*)(b->h_out + b->off_term); b->h_trunc = (uint8_t*)(b->h_out + b->off_trunc);
  b->d_obs = (float*)b->d_out; b->d_rew = (float*)(b->d_out + b->off_rew); b->d_term = (uint8_t*)(b->d_out + b->off_term); b->d_trunc = (uint8_t*)(b->d_out + b->off_trunc);
  CK(cudaMalloc(&v.final_ti, N * v.nti * 4)); CK(cudaMalloc(&v.final_tf, N * v.ntf * 4)); CK(cudaMalloc(&v.final_xpos, N * (size_t)dim[DD_nbody] * 3 * 4));
  CK(cudaMemset(v.final_ti, 0, N * v.nti * 4)); CK(cudaMemset(v.final_tf, 0, N * v.ntf * 4)); CK(cudaMemset(v.final_xpos, 0, N * (size_t)dim[DD_nbody] * 3 * 4));
  CK(cudaStreamCreateWithFlags(&b->own_stream, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&b->ev_order, cudaEventDisableTiming));
  { int rc = configure(b); if (rc) return rc; }
  guard.b = nullptr;
  *out = b;
  return B2_OK;
}
void b2_batch_destroy(B2Batch* b) {
  if (!b) return;
  cudaSetDevice(b->m->device);
  BatchView& v = b->v;
  cudaFree(v.wide); cudaFree(v.queue); cudaFree(v.cost); cudaFree(b->d_order);
  cudaFree(v.qpos); cudaFree(v.qvel); cudaFree(v.warm); cudaFree(v.qfrc_applied); cudaFree(v.ctrl); cudaFree(v.time);
  cudaFree(v.ti); cudaFree(v.tf); cudaFree(v.phase_cycles); cudaFree(v.counters); cudaFree(b->epstat); cudaFree(b->d_stats);
  cudaFree(v.final_ti); cudaFree(v.final_tf); cudaFree(v.final_xpos);
  cudaFreeHost(b->h_act); cudaFreeHost(b->h_out); cudaFree(b->d_act); cudaFree(b->d_out); cudaFree(b->d_inject);
  if (b->roll_exec) cudaGraphExecDestroy(b->roll_exec);
  if (b->ev_order) cudaEventDestroy(b->ev_order);
  if (b->own_stream) cudaStreamDestroy(b->own_stream);
  delete b;
}

int b2_dims(const B2Batch* b, int* o) {  /* o has 16 entries */
  if (!b || !o) return fail(B2_ERR_ARG, "null argument");
  const int* d = b->m->dm.dim;
  o[0] = d[DD_nq]; o[1] = d[DD_nv]; o[2] = d[DD_nu]; o[3] = d[DD_nbody]; o[4] = b->obs_dim; o[5] = b->act_dim; o[6] = b->n_envs;
  o[7] = b->nti; o[8] = b->ntf; o[9] = b->v.con_cap; o[10] = (int)b->smem; o[11] = b->v.envs_per_block; o[12] = b->v.row_cap; o[13] = b->m->dm.dim[DD_nM]; o[14] = b->v.arena_floats; o[15] = b->v.ws_floats * 4;
  return B2_OK;
}

int b2_reset(B2Batch* b, const uint8_t* mask_dev, const float* inject_dev, float* obs_dev, void* stream) {
  if (!b || !obs_dev) return fail(B2_ERR_ARG, "null argument");
  if (b->tp.task == TASK_NONE) return fail(B2_ERR_UNSUPPORTED, "batch has no task; use b2_set_state");
  b->v.reset_mask = mask_dev; b->v.obs = obs_dev; b->v.final_obs = nullptr; b->v.c_ncon = nullptr; b->v.xpos_out = nullptr;
  return launch(b, MODE_RESET, inject_dev, (cudaStream_t)stream);
}
int b2_step(B2Batch* b, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* term_dev, uint8_t* trunc_dev,
            float* final_obs_dev, void* stream) {
  if (!b || !act_dev || !obs_dev || !rew_dev || !term_dev || !trunc_dev) return fail(B2_ERR_ARG, "null argument");
  if (b->tp.task == TASK_NONE) return fail(B2_ERR_UNSUPPORTED, "batch has no task; use b2_physics_step");
  BatchView& v = b->v;
  v.action = act_dev; v.obs = obs_dev; v.reward = rew_dev; v.term = term_dev; v.trunc = trunc_dev; v.final_obs = final_obs_dev;
  v.reset_mask = nullptr; v.c_ncon = nullptr; v.xpos_out = nullptr;
  return launch(b, MODE_STEP, nullptr, (cudaStream_t)stream);
}
int b2_step_host(B2Batch* b, const float* act, float* obs, float* rew, uint8_t* term, uint8_t* trunc) {
  if (!b || !act || !obs || !rew || !term || !trunc) return fail(B2_ERR_ARG, "null argument");
  CK(cudaSetDevice(b->m->device));
  size_t N = b->n_envs; cudaStream_t s = b->own_stream;
  if (act != b->h_act) memcpy(b->h_act, act, N * b->act_dim * 4);          // callers holding b2_host_buffers() skip the staging copies
  CK(cudaMemcpyAsync(b->d_act, b->h_act, N * b->act_dim * 4, cudaMemcpyHostToDevice, s));
  int rc = b2_step(b, b->d_act, b->d_obs, b->d_rew, b->d_term, b->d_trunc, nullptr, s);
  if (rc) return rc;
  CK(cudaMemcpyAsync(b->h_out, b->d_out, b->out_bytes, cudaMemcpyDeviceToHost, s));     // obs | rew | term | trunc in one copy
  CK(cudaStreamSynchronize(s));
  b->pending = false;
  if (obs != b->h_obs) memcpy(obs, b->h_obs, N * b->obs_dim * 4);
  if (rew != b->h_rew) memcpy(rew, b->h_rew, N * 4);
  if (term != b->h_term) memcpy(term, b->h_term, N);
  if (trunc != b->h_trunc) memcpy(trunc, b->h_trunc, N);
  return B2_OK;
}
// T x env.step with the actions already on the device: the caller of env.step in test_parkour.py:45-47 / test_soccer.py:57-59 with
// the host taken out of the loop.  The T launches are captured once into a CUDA graph (kernel parameters carry the per-step
// slices of the caller's buffers) and replayed by one cudaGraphLaunch; the graph is re-captured when T or a buffer changes.
int b2_rollout(B2Batch* b, int T, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* term_dev, uint8_t* trunc_dev,
               float* final_obs_dev, void* stream) {
  if (!b || T <= 0 || !act_dev || !obs_dev || !rew_dev || !term_dev || !trunc_dev) return fail(B2_ERR_ARG, "bad argument");
  if (b->tp.task == TASK_NONE) return fail(B2_ERR_UNSUPPORTED, "batch has no task");
  CK(cudaSetDevice(b->m->device));
  cudaStream_t s = (cudaStream_t)stream; size_t N = b->n_envs;
  { int rc = order_after_last(b, s); if (rc) return rc; }
  const void* key[6] = {act_dev, obs_dev, rew_dev, term_dev, trunc_dev, final_obs_dev};
  if (!b->roll_exec || b->roll_T != T || memcmp(key, b->roll_key, sizeof(key)) != 0) {
    if (b->roll_exec) { cudaGraphExecDestroy(b->roll_exec); b->roll_exec = nullptr; }
    cudaStream_t cap = s ? s : b->own_stream;            // the legacy default stream cannot be captured
    cudaGraph_t g = nullptr;
    const unsigned long long launches_before = g_launches.load();
    CK(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    int rc = B2_OK;
    for (int t = 0; t < T && rc == B2_OK; t++)
      rc = b2_step(b, act_dev + (size_t)t * N * b->act_dim, obs_dev + (size_t)t * N * b->obs_dim, rew_dev + (size_t)t * N, term_dev + (size_t)t * N,
                   trunc_dev + (size_t)t * N, final_obs_dev ? final_obs_dev + (size_t)t * N * b->obs_dim : nullptr, cap);
    cudaError_t e = cudaStreamEndCapture(cap, &g);
    b->roll_launches = g_launches.load() - launches_before; g_launches = launches_before;      // counted per replay below, not per capture
    if (rc != B2_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(B2_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&b->roll_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) { b->roll_exec = nullptr; return fail(B2_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    b->roll_T = T; memcpy(b->roll_key, key, sizeof(key));
  }
  CK(cudaGraphLaunch(b->roll_exec, s));                // launching into the legacy stream is fine; only capturing on it is not
  g_launches += b->roll_launches;
  return B2_OK;
}
int b2_host_buffers(B2Batch* b, float** act, float** obs, float** rew, uint8_t** term, uint8_t** trunc) {
  if (!b) return fail(B2_ERR_ARG, "null argument");
  if (act) *act = b->h_act; if (obs) *obs = b->h_obs; if (rew) *rew = b->h_rew; if (term) *term = b->h_term; if (trunc) *trunc = b->h_trunc;
  return B2_OK;
}
int b2_reseed(B2Batch* b, uint64_t seed, void* stream) {
  if (!b) return fail(B2_ERR_ARG, "null argument");
  CK(cudaSetDevice(b->m->device));
  b->v.seed = seed;
  if (b->episode_slot >= 0) {      // episode counters restart, so reset(seed = s) always draws the same streams
    { int rc = order_after_last(b, (cudaStream_t)stream); if (rc) return rc; }
    CK(cudaMemset2DAsync(b->v.ti + b->episode_slot, (size_t)b->v.nti * 4, 0, 4, (size_t)b->n_envs, (cudaStream_t)stream));
  }
  return B2_OK;
}
int b2_caps(const B2Batch* b, int* o) {  /* o has 16 entries */
  if (!b || !o) return fail(B2_ERR_ARG, "null argument");
  memset(o, 0, 16 * sizeof(int));
  o[0] = b->ninj; o[1] = b->v.w_con_cap; o[2] = b->v.w_row_cap; o[3] = b->v.w_arena_floats; o[4] = (int)(b->v.wide_stride * 4 / 1024);
  o[5] = b->v.raw_cap; o[6] = b->v.act_cap; o[7] = b->v.wide ? 1 : 0; o[8] = b->episode_slot; o[9] = b->v.warm_once;
  return B2_OK;
}
int b2_get_final_state(B2Batch* b, int32_t* ti, float* tf, float* xpos, void* stream) {
  if (!b) return fail(B2_ERR_ARG, "null argument");
  CK(cudaSetDevice(b->m->device));
  size_t N = b->n_envs; cudaStream_t s = (cudaStream_t)stream;
  { int rc = order_after_last(b, s); if (rc) return rc; }
  if (ti && b->nti) CK(cudaMemcpyAsync(ti, b->v.final_ti, N * b->v.nti * 4, cudaMemcpyDeviceToDevice, s));
  if (tf && b->ntf) CK(cudaMemcpyAsync(tf, b->v.final_tf, N * b->v.ntf * 4, cudaMemcpyDeviceToDevice, s));
  if (xpos) CK(cudaMemcpyAsync(xpos, b->v.final_xpos, N * (size_t)b->m->dm.dim[DD_nbody] * 3 * 4, cudaMemcpyDeviceToDevice, s));
  return B2_OK;
}
int b2_physics_step(B2Batch* b, int nsub, void* stream) {
  if (!b || nsub < 0) return fail(B2_ERR_ARG, "bad argument");
  b->v.nsub = nsub; b->v.c_ncon = nullptr; b->v.xpos_out = nullptr;
  return launch(b, MODE_PHYS, nullptr, (cudaStream_t)stream);
}
int b2_forward(B2Batch* b, void* stream) {
  if (!b) return fail(B2_ERR_ARG, "null argument");
  b->v.c_ncon = nullptr; b->v.xpos_out = nullptr; b->v.forward_saves_warm = 1;
  int rc = launch(b, MODE_FORWARD, nullptr, (cudaStream_t)stream);
  b->v.forward_saves_warm = 0;
  return rc;
}

static int copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaStream_t s) {
  CK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToDevice, s));
  return B2_OK;
}
int b2_get_state(B2Batch* b, float* q, float* qv, float* c, float* wm, float* t, void* stream) {
  if (!b) return fail(B2_ERR_ARG, "null argument");
  CK(cudaSetDevice(b->m->device));
  { int rc = order_after_last(b, (cudaStream_t)stream); if (rc) return rc; }
  const int* d = b->m->dm.dim; BatchView& v = b->v; cudaStream_t s = (cudaStream_t)stream; size_t N = b->n_envs; int rc = 0;
  if (q) rc |= copy2d(q, d[DD_nq] * 4, v.qpos, v.nqp * 4, d[DD_nq] * 4, N, s);
  if (qv) rc |= copy2d(qv, d[DD_nv] * 4, v.qvel, v.nvp * 4, d[DD_nv] * 4, N, s);
  if (c && d[DD_nu]) rc |= copy2d(c, d[DD_nu] * 4, v.ctrl, v.nup * 4, d[DD_nu] * 4, N, s);
  if (wm) rc |= copy2d(wm, d[DD_nv] * 4, v.warm, v.nvp * 4, d[DD_nv] * 4, N, s);
  if (t) CK(cudaMemcpyAsync(t, v.time, N * 4, cudaMemcpyDeviceToDevice, s));
  return rc ? B2_ERR_CUDA : B2_OK;
}
int b2_set_state(B2Batch* b, const float* q, const float* qv, const float* c, const float* wm, const float* t, void* stream) {
  if (!b) return fail(B2_ERR_ARG, "null argument");
  CK(cudaSetDevice(b->m->device));
  { int rc = order_after_last(b, (cudaStream_t)stream); if (rc) return rc; }
  const int* d = b->m->dm.dim; BatchView& v = b->v; cudaStream_t s = (cudaStream_t)stream; size_t N = b->n_envs; int rc = 0;
  if (q) rc |= copy2d(v.qpos, v.nqp * 4, q, d[DD_nq] * 4, d[DD_nq] * 4, N, s);
  if (qv) rc |= copy2d(v.qvel, v.nvp * 4, qv, d[DD_nv] * 4, d[DD_nv] * 4, N, s);
  if (c && d[DD_nu]) rc |= copy2d(v.ctrl, v.nup * 4, c, d[DD_nu] * 4, d[DD_nu] * 4, N, s);
  if (wm) rc |= copy2d(v.warm, v.nvp * 4, wm, d[DD_nv] * 4, d[DD_nv] * 4, N, s);
  if (t) CK(cudaMemcpyAsync(v.time, t, N * 4, cudaMemcpyDeviceToDevice, s));
  return rc ? B2_ERR_CUDA : B2_OK;
}
int b2_get_task_state(B2Batch* b, int32_t* ti, float* tf, void* stream) {
  if (!b) return fail(B2_ERR_ARG, "null argument");
  CK(cudaSetDevice(b->m->device));
  { int rc = order_after_last(b, (cudaStream_t)stream); if (rc) return rc; }
  size_t N = b->n_envs; cudaStream_t s = (cudaStream_t)stream;
  if (ti && b->nti) CK(cudaMemcpyAsync(ti, b->v.ti, N * b->v.nti * 4, cudaMemcpyDeviceToDevice, s));
  if (tf && b->ntf) CK(cudaMemcpyAsync(tf, b->v.tf, N * b->v.ntf * 4, cudaMemcpyDeviceToDevice, s));
  return B2_OK;
}
int b2_set_task_state(B2Batch* b, const int32_t* ti, const float* tf, void* stream) {
  if (!b) return fail(B2_ERR_ARG, "null argument");
  CK(cudaSetDevice(b->m->device));
  { int rc = order_after_last(b, (cudaStream_t)stream); if (rc) return rc; }
  size_t N = b->n_envs; cudaStream_t s = (cudaStream_t)stream;
  if (ti && b->nti) CK(cudaMemcpyAsync(b->v.ti, ti, N * b->v.nti * 4, cudaMemcpyDeviceToDevice, s));
  if (tf && b->ntf) CK(cudaMemcpyAsync(b->v.tf, tf, N * b->v.ntf * 4, cudaMemcpyDeviceToDevice, s));
  return B2_OK;
}
int b2_get_contacts(B2Batch* b, int32_t* ncon, int32_t* geom, float* dist, int cap, void* stream) {
  if (!b || !ncon || !geom || !dist || cap <= 0) return fail(B2_ERR_ARG, "bad argument");
  b->v.c_ncon = ncon; b->v.c_geom = geom; b->v.c_dist = dist; b->v.c_cap = cap; b->v.xpos_out = nullptr;
  int rc = launch(b, MODE_FORWARD, nullptr, (cudaStream_t)stream);
  b->v.c_ncon = nullptr;
  return rc;
}
int b2_get_xpos(B2Batch* b, float* xpos, void* stream) {
  if (!b || !xpos) return fail(B2_ERR_ARG, "bad argument");
  b->v.c_ncon = nullptr; b->v.xpos_out = xpos;
  int rc = launch(b, MODE_FORWARD, nullptr, (cudaStream_t)stream);
  b->v.xpos_out = nullptr;
  return rc;
}
int b2_debug_forward(B2Batch* b, float* out_dev, int n_per_env, void* stream) {
  if (!b || !out_dev || n_per_env <= 0) return fail(B2_ERR_ARG, "bad argument");
  b->v.c_ncon = nullptr; b->v.xpos_out = nullptr; b->v.debug_out = out_dev; b->v.debug_n = n_per_env;
  int rc = launch(b, MODE_FORWARD, nullptr, (cudaStream_t)stream);
  b->v.debug_out = nullptr;
  return rc;
}
int b2_stats(B2Batch* b, double* out, void* stream) {
  if (!b || !out) return fail(B2_ERR_ARG, "null argument");
  CK(cudaSetDevice(b->m->device));
  { int rc = order_after_last(b, (cudaStream_t)stream); if (rc) return rc; }
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemsetAsync(out, 0, 16 * 8, s));
  int blocks = (b->n_envs + 255) / 256; if (blocks > 148) blocks = 148;
  b2_stats_kernel<<<blocks, 256, 0, s>>>(b->v.counters, b->epstat, b->n_envs, out);
  g_launches++;
  CK(cudaGetLastError());
  return B2_OK;
}

}  // extern "C"
