/* b2env.h -- C-ABI of the B200 batched physics-and-task engine (libb2env.so).
 *
 * The reference (hasnainfarid/Mujoco_Gymnasium_Environments) has no FFI of its own: its hot path is Python calling
 * the `mujoco` bindings.  Each entry point below names the reference call it stands in for, so that a maintainer can
 * bind it with ctypes from the reference's env classes (INTEGRATION.md shows the stub).
 *
 * Conventions: every function returns 0 on success or a negative code and never throws across the ABI;
 * b2_last_error() gives the message.  Pointers named *_dev are CUDA device pointers owned by the caller (e.g. torch
 * tensors); the library borrows them for the duration of the stream-ordered call.  `stream` is a cudaStream_t passed
 * as void* (NULL = default stream).  One host thread per B2Batch.  No global mutable state besides the error string.
 */
#ifndef B2ENV_H
#define B2ENV_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct B2Model B2Model;
typedef struct B2Batch B2Batch;

enum { B2_OK = 0, B2_ERR_ARG = -1, B2_ERR_CUDA = -2, B2_ERR_LAYOUT = -3, B2_ERR_UNSUPPORTED = -4 };
enum { B2_TASK_NONE = 0, B2_TASK_QUADRUPED_PARKOUR = 1, B2_TASK_HUMANOID_DANCING = 2, B2_TASK_HUMANOID_SOCCER = 3, B2_TASK_BIPEDAL_RESCUE = 4, B2_TASK_HUMANOID_CONSTRUCTION = 5, B2_TASK_HUMANOID_MARTIAL_ARTS = 6, B2_TASK_ROBOTIC_ARM_ASSEMBLY = 7 };

/* Replaces mujoco.MjModel.from_xml_string(...) (quadruped_parkour_env/parkour_env.py:100 and the six sibling call
 * sites): takes the packed device tables produced by the Python-side compiler (device_pack.pack_device_model) and
 * uploads them (fp64 -> fp32) to `device`. */
int b2_model_create(const int32_t* ints, int n_ints, const double* flts, int n_flts, int device, B2Model** out);
void b2_model_destroy(B2Model* m);

/* Task description resolved by the host from names (the reference's mj_name2id lookups, parkour_env.py:181-232). */
typedef struct B2TaskDesc {
  int task;            /* B2_TASK_* */
  int ids[16];
  float act_lo[40], act_hi[40];
  int aux_i[64];       /* task-specific integer table (arm: per-geom component index from the reference's geom-name matching) */
  float aux_f[32];     /* task-specific float table (arm: component targets, assembly_env.py:65-75, and the ee_site offset) */
} B2TaskDesc;

/* Fixed capacities of the per-env on-chip buffers; 0 = library default.  A forward pass whose contacts / rows exceed
 * them runs in the wide tier (per-env global workspace, same arithmetic); only what exceeds the wide capacities too
 * (256 contacts, 1024 rows) is dropped and counted (b2_stats), never silently. */
typedef struct B2BatchOpts {
  int envs_per_block;  /* warps (= envs) per CTA sharing one staged model copy */
  int arena_floats;    /* per-env shared-memory arena holding J and the packed A of every island */
  int con_cap;         /* contact buffer capacity per env */
  int row_cap;         /* constraint-row capacity per env */
  int warps_per_env;   /* 3: warps cooperating on one env (islands / contact chain in parallel); 0 = default */
  int disable_wide;    /* non-zero: no global spill workspace; what exceeds the on-chip capacities is dropped and counted (A/B tests) */
  int warmstart_once_per_step; /* 0 (default): qacc_warmstart is saved at the end of every forward pass, as MuJoCo 3.x's mj_fwdConstraint does
                                  (RK4 stages 2-4 start from the previous stage); non-zero: once per mj_step (the round-1 reading) */
  int fifo_queue;      /* non-zero: envs are stepped in index order; default: longest-last-step first (a long env pulled last is a tail) */
} B2BatchOpts;

/* Replaces mujoco.MjData(model) for n_envs lock-stepped environments (parkour_env.py:54).  env_offset is the global
 * index of this shard's env 0 so RNG streams do not depend on the number of GPUs.  opts may be NULL. */
int b2_batch_create(B2Model* m, const B2TaskDesc* task, int n_envs, uint64_t seed, int env_offset, const B2BatchOpts* opts,
                    B2Batch** out);
void b2_batch_destroy(B2Batch* b);

/* dims (16 ints): [nq, nv, nu, nbody, obs_dim, act_dim, n_envs, nti, ntf, con_cap, smem_bytes_per_cta, envs_per_block,
 * row_cap, nM, arena_floats, smem_bytes_per_env] */
int b2_dims(const B2Batch* b, int* out16);

/* Env.reset() for the envs whose mask byte is non-zero (NULL = all): mj_resetData + task reset + randomisation +
 * settle steps + first observation (parkour_env.py:314-354).  inject_dev (nullable, [n_envs][4]) overrides the
 * random draws so parity tests can reproduce the oracle's reset exactly. */
int b2_reset(B2Batch* b, const uint8_t* mask_dev, const float* inject_dev, float* obs_dev, void* stream);

/* Env.step(action) for all envs (parkour_env.py:356-394): clip, ctrl, frame_skip x mj_step, obs, reward,
 * terminated, truncated, and same-step auto-reset; final_obs_dev (nullable) receives the pre-reset observation. */
int b2_step(B2Batch* b, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* term_dev, uint8_t* trunc_dev,
            float* final_obs_dev, void* stream);

/* Same call with HOST buffers (pageable or pinned): one H2D copy of the actions, one launch, one D2H copy of the packed
 * [obs | rew | term | trunc] block, synchronises.  This is the end-to-end path a CPU-side caller of env.step would use
 * (the host loop of quadruped_parkour_env/test_parkour.py:45-47). */
int b2_step_host(B2Batch* b, const float* act, float* obs, float* rew, uint8_t* term, uint8_t* trunc);

/* T consecutive Env.step calls with the host out of the loop (the caller of env.step: quadruped_parkour_env/test_parkour.py:45-47,
 * humanoid_soccer_env/test_soccer.py:57-59, given T pre-computed actions): act_dev [T][n_envs][act_dim] in, obs_dev
 * [T][n_envs][obs_dim], rew_dev / term_dev / trunc_dev [T][n_envs] (and final_obs_dev, nullable, like obs_dev) out.  The T
 * launches are one CUDA graph, captured on first use and replayed while T and the buffers stay the same.  b2_step itself is
 * capturable too (no attribute / device calls on its launch path), e.g. inside a torch.cuda.graph region with a policy. */
int b2_rollout(B2Batch* b, int T, const float* act_dev, float* obs_dev, float* rew_dev, uint8_t* term_dev, uint8_t* trunc_dev,
               float* final_obs_dev, void* stream);

/* The library's own pinned staging buffers for b2_step_host ([n_envs][act_dim], [n_envs][obs_dim], [n_envs] ...): a caller
 * that fills / reads these directly and passes them to b2_step_host pays no host-side staging copy.  Any pointer may be NULL. */
int b2_host_buffers(B2Batch* b, float** act, float** obs, float** rew, uint8_t** term, uint8_t** trunc);

/* Env.reset(seed=...) (parkour_env.py:314-322, super().reset(seed=seed) reseeds np_random): replaces the RNG seed and restarts
 * the per-env episode counters, so b2_reseed(s) + b2_reset always produces the same initial states. */
int b2_reseed(B2Batch* b, uint64_t seed, void* stream);

/* caps (16 ints): [n_inject (floats per env of b2_reset's inject_dev), wide_con_cap, wide_row_cap, wide_arena_floats,
 * wide_workspace_KiB_per_env, raw_cap, act_cap, wide_enabled, episode_slot, warmstart_once_per_step, 0...] */
int b2_caps(const B2Batch* b, int* out16);

/* Task state / data.xpos of the episode that ended in an env's last terminal step, snapshotted before the same-step
 * auto-reset ([n_envs][nti], [n_envs][ntf], [n_envs][nbody][3]; rows of envs that never finished an episode are zero): what
 * the reference's terminal-step `info` is built from (parkour_env.py:797-813).  Any pointer may be NULL. */
int b2_get_final_state(B2Batch* b, int32_t* ti_dev, float* tf_dev, float* xpos_dev, void* stream);

/* nsub x mujoco.mj_step(model, data) on the raw state, no task logic (parkour_env.py:348,368). */
int b2_physics_step(B2Batch* b, int nsub, void* stream);
/* mujoco.mj_forward(model, data) (humanoid_martial_arts_env/martial_arts_env.py:481): refreshes contacts / xpos and,
 * like mj_fwdConstraint, leaves qacc in qacc_warmstart (b2_get_contacts / b2_get_xpos / b2_debug_forward do not). */
int b2_forward(B2Batch* b, void* stream);

/* data.qpos / qvel / ctrl / qacc_warmstart / time access (dense row-major [n_envs][dim] fp32 device arrays; any
 * pointer may be NULL).  Doubles as checkpoint/restore. */
int b2_get_state(B2Batch* b, float* qpos_dev, float* qvel_dev, float* ctrl_dev, float* warm_dev, float* time_dev, void* stream);
int b2_set_state(B2Batch* b, const float* qpos_dev, const float* qvel_dev, const float* ctrl_dev, const float* warm_dev,
                 const float* time_dev, void* stream);
int b2_get_task_state(B2Batch* b, int32_t* ti_dev, float* tf_dev, void* stream);
int b2_set_task_state(B2Batch* b, const int32_t* ti_dev, const float* tf_dev, void* stream);

/* data.ncon / data.contact[i].geom1,geom2,dist of the last forward pass: ncon_dev [n_envs], geom_dev
 * [n_envs][cap][2] (model geom ids), dist_dev [n_envs][cap]. */
int b2_get_contacts(B2Batch* b, int32_t* ncon_dev, int32_t* geom_dev, float* dist_dev, int cap, void* stream);
/* data.xpos of the last forward pass, [n_envs][nbody][3]. */
int b2_get_xpos(B2Batch* b, float* xpos_dev, void* stream);

/* Bring-up / test hook: mj_forward, then dump per env [qfrc_smooth|qacc_smooth|qfrc_constraint|qacc (nv each)|
 * M (nM, sparse)|ncon nefc solver_iter 0|efc_force|efc_b|efc_R|efc_pos (row_cap each, island-major row order)]. */
int b2_debug_forward(B2Batch* b, float* out_dev, int n_per_env, void* stream);

/* Episode statistics and engine counters summed over this shard into out_dev[16] (fp64), ready for an NCCL
 * all-reduce: [episodes, return_sum, length_sum, nan_resets, contacts_dropped, rows_dropped, arena_overflows,
 * solver_iters, substeps, newton_iteration_caps, wide_passes (forward passes run in the wide tier), wide_passes_rows (of those: chosen because of rows /
 * arena space rather than the contact count), 0...]. */
int b2_stats(B2Batch* b, double* out_dev16, void* stream);

/* kernels launched by this library since load (the bench's gpu_launches claim) */
unsigned long long b2_launch_count(void);
const char* b2_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
