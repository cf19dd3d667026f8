// b2_engine.cuh -- lock-step mj_step, one environment per warp, E environments per CTA, sm_100a.
//
// Replaces the arithmetic behind the reference's `mujoco.mj_step(model, data)` call sites
// (quadruped_parkour_env/parkour_env.py:348,368 and siblings) for the feature
// subset in SURVEY App. B.  Design:
//   * the model tables are staged once per CTA into shared memory with one TMA bulk copy per buffer
//     (cp.async.bulk + mbarrier) and shared by the CTA's E warps;
//   * each warp owns one env: every per-env intermediate (frames, spatial inertias, sparse M and its L'DL factor,
//     contacts, J, A) lives in that warp's slice of shared memory for the whole env-step; HBM is touched only to
//     load/store qpos/qvel/warmstart/ctrl/task state; all synchronisation is __syncwarp (no CTA barriers after
//     model staging), so warps never wait for each other (round-1 ncu: the CTA-per-env variant spent ~89% of issue
//     slots idle at __syncthreads, profiles/r01_cta_per_env.txt);
//   * tree passes are level-synchronous over bodies / dofs with lanes as bodies / dofs; the constraint problem is
//     split into islands (kinematic trees that can exchange contact forces); each island's A = J M^-1 J' + R is kept
//     symmetric-packed in shared memory and its projected Gauss-Seidel sweep keeps forces and residuals in registers
//     (one shuffle broadcast + S FMAs per row).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2_device_layout.h"
#include "b2_math.cuh"
#include "b2_collide.cuh"

namespace b2 {

#define B2_MINVAL 1e-15f
#define B2_MAXVAL 1e10f
#define B2_MINIMP 0.0001f
#define B2_MAXIMP 0.9999f
#define B2_MAX_ISLANDS 16
#define B2_CON_STRIDE 16   // floats per contact record
#define B2_ISLAND_ROWS 128  // rows per island: each lane of the sweeping warp owns 4 consecutive rows in registers
#define B2_FULL 0xffffffffu
#define B2_WREC 40          // floats of a wide PGS block's record: 24 (chain coefficients) + 16 (look-ahead tile)

struct DevModel {
  const int* ints; const float* flts;
  int n_ints, n_flts;
  int n_ints_staged;               // ints copied into shared memory (all of them, or up to the pair tables when cold)
  int ioff[B2DEV_N_INT_FIELDS];
  int foff[B2DEV_N_FLT_FIELDS];
  int dim[DD_COUNT];
  float opt[DO_COUNT];
};

enum { CTR_NAN_RESET = 0, CTR_CON_DROPPED = 1, CTR_ROW_DROPPED = 2, CTR_ARENA_OVERFLOW = 3, CTR_EPISODES = 4,
       CTR_SOLVER_ITERS = 5, CTR_SUBSTEPS = 6, CTR_ARENA_SPILL = 7, CTR_WIDE = 8, CTR_WIDE_ROWS = 9, CTR_COUNT = 10 };

// -------------------------------------------------------------------------------------------- workspace (per warp)
// All per-env intermediates live in the warp's slice of dynamic shared memory.  The slice is addressed as
// b2_smem[warp_base + offset + i] with the offsets below held in the kernel-parameter constant bank, so every access
// compiles to LDS/STS with no pointer registers (round-1 ncu: a struct of 48 generic pointers spilled to local memory
// and turned ~2/3 of the shared-memory traffic into generic LD/ST, profiles/r01_warp_per_env_generic.txt).
extern __shared__ __align__(128) float b2_smem[];

#define B2_WS_FLOAT_FIELDS(X) X(qpos) X(qvel) X(warm) X(ctrl) X(qapp) X(xpos) X(xmat) X(cdof) X(rootcom) \
  X(xquat) X(xipos) X(cvel) X(cacc) X(cinert) X(M) X(LD) X(invD) X(qfs) X(qas) X(qfc) X(qacc) X(tmp) X(con) \
  X(row_R) X(row_b) X(row_f) X(row_res) X(arena) X(red) X(time) X(tf) X(act) X(rk_q0) X(rk_v0) X(rk_sv) X(rk_sa) X(xfrc) X(LD2) X(invD2)
#define B2_WS_INT_FIELDS(X) X(lim_row) X(con_row) X(row_info) X(isl_n) X(isl_nl) X(isl_warp) X(isl_nd) X(isl_col0) X(tree_isl) X(tree_col0) X(dof_col) X(col_dof) X(isl_adr) X(isl_J) X(isl_A) X(isl_ldj) X(misc) X(ti)

struct WsOff {
#define X(n) int n;
  B2_WS_FLOAT_FIELDS(X) B2_WS_INT_FIELDS(X)
#undef X
};
// offsets (floats) of the spillable fields inside one env's wide workspace
struct WideOff { int con, con_row, row_info, row_R, row_b, row_f, row_res, arena; };
__host__ inline long long wide_layout(int w_con_cap, int w_row_cap, int w_arena_floats, WideOff* o) {
  long long off = 0;
  auto take = [&](int n) { long long r = off; off += (n + 3) & ~3; return (int)r; };
  WideOff t;
  t.con = take(w_con_cap * 16); t.con_row = take(w_con_cap); t.row_info = take(w_row_cap); t.row_R = take(w_row_cap);
  t.row_b = take(w_row_cap); t.row_f = take(w_row_cap); t.row_res = take(w_row_cap); t.arena = take(w_arena_floats);
  if (o) *o = t;
  return (off + 31) & ~31ll;
}
enum { MISC_NCON = 0, MISC_NEFC = 1, MISC_FLAG = 2, MISC_ITERS = 3, MISC_ARENA_USED = 4, MISC_DONE = 5, MISC_NISL = 6, MISC_WIDE = 7, MISC_WSCR = 8, MISC_ENV = 9, MISC_WRING = 10, MISC_WRB = 11, MISC_COUNT = 12 };

__host__ __device__ inline int r4(int n) { return (n + 3) & ~3; }
// model tables occupy the first model_floats of shared memory (ints first, then floats), padded to 32 floats
__host__ __device__ inline int model_smem_floats(int n_ints_staged, int n_flts) { return (r4(n_ints_staged) + r4(n_flts) + 31) & ~31; }

// per-warp carve-up in floats; returns the slice size (multiple of 32 floats)
__host__ inline int ws_layout(const int* dim, int con_cap, int row_cap, int arena_floats, int nti, int ntf, WsOff* o) {
  int off = 0;
  auto take = [&](int n_words) { int r = off; off += r4(n_words); return r; };
  int nq = dim[DD_nq], nv = dim[DD_nv], nu = dim[DD_nu], nb = dim[DD_nbody], nM = dim[DD_nM];
  int nroot = dim[DD_nroot], nlim = dim[DD_nlim];
  WsOff t;
  t.qpos = take(nq); t.qvel = take(nv); t.warm = take(nv); t.ctrl = take(nu > 0 ? nu : 1); t.qapp = take(nv);
  t.xpos = take(3 * nb);
  // contiguous block [xmat .. rootcom]: dead once the J rows are filled, reused as scratch by the A build; xquat and
  // rootcom sit at its end so a task that reads them after the step (keep_frames) only shortens the reusable part
  t.xmat = take(9 * nb); t.cdof = take(6 * nv); t.xipos = take(3 * nb); t.cvel = take(6 * nb); t.cacc = take(6 * nb);
  t.cinert = take(10 * nb); t.xquat = take(4 * nb); t.rootcom = take(3 * (nroot > 0 ? nroot : 1));
  t.M = take(nM); t.LD = take(nM); t.invD = take(nv); t.qfs = take(nv); t.qas = take(nv); t.qfc = take(nv);
  t.qacc = take(nv); t.tmp = take(nv);
  t.con = take(con_cap * B2_CON_STRIDE); t.lim_row = take(2 * (nlim > 0 ? nlim : 1)); t.con_row = take(con_cap);
  t.row_info = take(row_cap); t.row_R = take(row_cap); t.row_b = take(row_cap); t.row_f = take(row_cap); t.row_res = take(row_cap);
  t.isl_n = take(B2_MAX_ISLANDS); t.isl_nl = take(B2_MAX_ISLANDS); t.isl_warp = take(B2_MAX_ISLANDS); t.isl_nd = take(B2_MAX_ISLANDS);
  t.isl_col0 = take(B2_MAX_ISLANDS + 4); t.tree_isl = take(B2_MAX_ISLANDS); t.tree_col0 = take(B2_MAX_ISLANDS); t.dof_col = take(nv); t.col_dof = take(nv);
  t.isl_adr = take(B2_MAX_ISLANDS + 4); t.isl_J = take(B2_MAX_ISLANDS);
  t.isl_A = take(B2_MAX_ISLANDS); t.isl_ldj = take(B2_MAX_ISLANDS);
  t.red = take(16); t.misc = take(MISC_COUNT); t.time = take(4); t.ti = take(nti > 0 ? nti : 1); t.tf = take(ntf > 0 ? ntf : 1);
  t.act = take(40);
  bool rk4 = dim[DD_integrator] == 1;
  t.rk_q0 = take(rk4 ? nq : 0); t.rk_v0 = take(rk4 ? nv : 0); t.rk_sv = take(rk4 ? nv : 0); t.rk_sa = take(rk4 ? nv : 0);
  t.xfrc = take(8);
  t.LD2 = take(rk4 ? 0 : nM); t.invD2 = take(rk4 ? 0 : nv);      // factor of M + h D (Euler), built beside the factor of M
  t.arena = take(arena_floats);
  if (o) *o = t;
  return (off + 31) & ~31;
}
struct BatchView {
  int n_envs; int first_env;                         // this launch steps envs [first_env, first_env + n_envs)
  float *qpos, *qvel, *warm, *ctrl, *qfrc_applied, *time;
  int nqp, nvp, nup;               // row pitches (floats)
  int *ti; float *tf; int nti, ntf;  // per-env task state
  float *obs, *final_obs, *reward; uint8_t *term, *trunc; int obs_dim;
  int* final_ti; float* final_tf; float* final_xpos; // task state / xpos of the episode that ended in this step (before the in-kernel reset)
  const float* action; int act_dim;
  const uint8_t* reset_mask;
  int *c_ncon, *c_geom; float* c_dist; int c_cap;   // optional contact export
  float* xpos_out;                                   // optional [N][nbody*3] export of the last forward pass
  float* debug_out; int debug_n;                     // optional [N][debug_n] dump of solver intermediates (bring-up / tests)
  unsigned long long* counters;                      // [N][CTR_COUNT]
  int* queue;                                        // [0] next env of this launch, [1] teams that have left (work queue of the persistent CTAs)
  unsigned* cost;                                    // [N] device cycles (>> 8) each env's last control step took
  const int* order;                                  // [N] queue position -> env, longest-last-step first (nullptr: identity)
  int lockstep;                                      // the teams of a CTA start every forward pass together (CTA barrier): they then run the same code at the same time and share instruction-cache lines
  unsigned long long* phase_cycles;                  // [16] per-phase clock64 sums (only with -DB2_PHASE_TIMING)
  unsigned long long seed;
  int arena_floats, con_cap, row_cap;
  int act_cap, raw_cap;                              // narrow phase: active-pair list entries, raw contact slots
  int nsub;                                          // physics sub-steps for MODE_PHYS
  int env_offset;                                    // global index of env 0 (multi-GPU sharding keeps RNG streams fixed)
  int envs_per_block; int ws_floats; int model_floats;   // shared-memory slices, in floats
  int keep_frames;                                   // the task reads xquat / subtree com of the last forward pass
  int inject_stride;                                 // floats per env of the injected reset draws
  int xfrc_body;                                     // body whose xfrc_applied the task drives (-1: none); value in ws.xfrc[0..5]
  int warm_once;                                     // 0 (default, MuJoCo 3.x): qacc_warmstart <- qacc at the end of every forward pass; 1: once per mj_step
  int forward_saves_warm;                            // MODE_FORWARD writes the warm start back (b2_forward = mj_forward); the read-only exports do not
  // wide tier: per-env spill workspace in global memory (L2-resident while in use).  A forward pass whose contacts / rows /
  // J do not fit the on-chip capacities keeps its contact records, row arrays and J there instead of dropping anything.
  float* wide; long long wide_stride;                // [n_envs][wide_stride] floats (nullptr: tier disabled)
  int w_con_cap, w_row_cap, w_arena_floats;
  WideOff woff;
  WsOff off;
};

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_PHYS = 2, MODE_FORWARD = 3 };

// floats of the contiguous dead block [xmat .. rootcom] the A build may overwrite
__host__ __device__ inline int dead_block_floats(const int* dim, int keep_frames) {
  int nb = dim[DD_nbody], nv = dim[DD_nv], nroot = dim[DD_nroot];
  if (keep_frames >= 2) return 0;      // the task reads xipos/xmat too: nothing of the block may be overwritten
  int n = r4(9 * nb) + r4(6 * nv) + r4(3 * nb) + r4(6 * nb) + r4(6 * nb) + r4(10 * nb);
  if (!keep_frames) n += r4(4 * nb) + r4(3 * (nroot > 0 ? nroot : 1));
  return n;
}

// -------------------------------------------------------------------------------------------- engine (one warp)
// One env is stepped by a team of W warps (W = 1: one warp per env).  Warp-synchronous phases (tree passes, row
// compaction) run on one warp of the team; the parallel phases (J fill, A build, PGS sweeps) split islands / rows
// over the team's warps; team_sync() is a named barrier private to the team.
// HOIST: a warp sweeping a single island keeps its PGS rows in registers across iterations; COOPMIN: islands with more
// rows than this are built (A = J M^-1 J') by the whole team.  Both are per-task tuning knobs (A/B-measured on B200).
// DYN: islands are re-formed every forward pass from the active contacts (needed when moving trees can touch each other);
// otherwise they are the model's static islands, set up once per launch.
// SOLVER: 0 the task's model uses PGS, 2 Newton (the other solver is not compiled into that kernel), -1 decided at run time
// from the model (physics-only batches).
// C6: the model has condim-6 pairs (10-row pyramids: two tangential, one torsional, two rolling directions); otherwise every
// contact is a 4-row pyramid and the row bookkeeping uses the cheaper fixed-size arithmetic.
// TEAMND: Newton islands with more dofs than this are solved by the whole team, smaller ones by one warp each.
// CVX: the model has candidate pairs that take the convex (MPR) path.
template <int W, bool HOIST = true, int COOPMIN = 32, bool COLD = false, bool DYN = true, int SOLVER = -1, bool C6 = false, int TEAMND = 16, bool CVX = true>
struct Engine {
  const DevModel& P;
  const BatchView& B;
  int wb;      // team's slice base in b2_smem (floats)
  int lane;    // lane in warp
  int wl;      // warp in team
  int tl;      // thread in team
  int barid;   // named barrier of the team (1..15)
  int env;     // env this team steps (row of the wide workspace)
  bool wide;   // this forward pass keeps its contacts / rows / J in the wide workspace (team-uniform, refreshed from MISC_WIDE)
  static constexpr int TEAM = 32 * W;
  static constexpr int RING_MIN = 3, RING_MAX = 16;      // stages of the wide PGS sweep's bulk-copy ring (depth chosen per pass from the free arena)

  __device__ Engine(const DevModel& p, const BatchView& b, int team_base, int team_in_block, int env_) : P(p), B(b), wb(team_base), env(env_), wide(false) {
    tl = threadIdx.x % TEAM; lane = tl & 31; wl = tl >> 5; barid = 1 + team_in_block;
  }
  __device__ __forceinline__ void team_sync() const {
    if (W == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(barid), "n"(TEAM) : "memory");
  }

#define X(n) __device__ __forceinline__ float* p_##n() const { return b2_smem + wb + B.off.n; }
  B2_WS_FLOAT_FIELDS(X)
#undef X
#define X(n) __device__ __forceinline__ int* p_##n() const { return (int*)(b2_smem + wb + B.off.n); }
  B2_WS_INT_FIELDS(X)
#undef X
  __device__ __forceinline__ const int* I(int f) const { return (const int*)b2_smem + P.ioff[f]; }
  __device__ __forceinline__ const float* F(int f) const { return b2_smem + r4(P.n_ints_staged) + P.foff[f]; }
  // candidate-pair tables: shared memory, or global memory (read-only, L2-resident) when the task leaves them cold
  __device__ __forceinline__ const int* PI(int f) const { return COLD ? P.ints + P.ioff[f] : (const int*)b2_smem + P.ioff[f]; }
  __device__ __forceinline__ int dim(int k) const { return P.dim[k]; }
  __device__ __forceinline__ bool newton() const { return SOLVER < 0 ? P.dim[DD_solver] == 2 : SOLVER == 2; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ int conCap() const { return wide ? B.w_con_cap : B.con_cap; }
  __device__ __forceinline__ int rowCap() const { return wide ? B.w_row_cap : B.row_cap; }
  __device__ __forceinline__ int arenaFloats() const { return B.arena_floats; }
  // spillable fields: shared memory in the fast tier, the env's wide workspace otherwise.  x_*() select at run time (generic
  // pointers; contact chain, task epilogues), xs_*<WD>() at compile time (row fill, solvers).
  __device__ __forceinline__ float* wbase() const { return B.wide + (size_t)env * (size_t)B.wide_stride; }
#define XF(n) __device__ __forceinline__ float* x_##n() const { return wide ? wbase() + B.woff.n : b2_smem + wb + B.off.n; } \
  template <bool WD> __device__ __forceinline__ float* xs_##n() const { return WD ? wbase() + B.woff.n : b2_smem + wb + B.off.n; }
#define XI(n) __device__ __forceinline__ int* x_##n() const { return (int*)(wide ? wbase() + B.woff.n : b2_smem + wb + B.off.n); } \
  template <bool WD> __device__ __forceinline__ int* xs_##n() const { return (int*)(WD ? wbase() + B.woff.n : b2_smem + wb + B.off.n); }
  XF(con) XF(row_R) XF(row_b) XF(row_f) XF(row_res) XI(con_row) XI(row_info)
#undef XF
#undef XI
  __device__ __forceinline__ void refresh_wide() { wide = p_misc()[MISC_WIDE] != 0; }
  // J of island k: fast tier / Newton: plain rows (blk = 4 ldj, row i at i * ldj); wide PGS: blocks of four rows
  // [J rows | B = M^-1 J' rows | record], ldj = r4(nd), so one aligned copy brings everything a 4-row sweep step needs
  __device__ __forceinline__ int jblk(int ldj) const { return (wide && !newton()) ? 8 * ldj + B2_WREC : 4 * ldj; }
  __device__ __forceinline__ float* x_J(int k) const { return (wide ? wbase() + B.woff.arena : p_arena()) + p_isl_J()[k]; }
  template <bool WD> __device__ __forceinline__ float* xs_J(int k) const { return (WD ? wbase() + B.woff.arena : p_arena()) + p_isl_J()[k]; }
  __device__ __forceinline__ static int jrow(int i, int blk, int ldj) { return (i >> 2) * blk + (i & 3) * ldj; }
  template <bool WD> __device__ __forceinline__ static int jr(int i, int blk, int ldj) { return WD ? jrow(i, blk, ldj) : i * ldj; }

  // ---- B.1 kinematics: level-synchronous over bodies, lanes = bodies of one depth level
  __device__ void kinematics() {
    const int* parent = I(DI_body_parentid); const int* jadr = I(DI_body_jntadr); const int* jnum = I(DI_body_jntnum);
    const int* jtype = I(DI_jnt_type); const int* jq = I(DI_jnt_qposadr); const int* jd = I(DI_jnt_dofadr);
    const int* ladr = I(DI_level_adr); const int* lbody = I(DI_level_body);
    const float* bpos = F(DF_body_pos); const float* bquat = F(DF_body_quat); const float* bipos = F(DF_body_ipos);
    const float* jpos = F(DF_jnt_pos); const float* jaxis = F(DF_jnt_axis); const float* q0 = F(DF_qpos0);
    if (lane == 0) {
      st3(p_xpos(), v3(0, 0, 0)); Q4 qi; qi.w = 1; qi.x = qi.y = qi.z = 0; stq(p_xquat(), qi); quat2mat(p_xmat(), qi);
      st3(p_xipos(), v3(0, 0, 0));
      for (int k = 0; k < 6; k++) p_cvel()[k] = 0.f;
      for (int k = 0; k < 10; k++) p_cinert()[k] = 0.f;
      p_cacc()[0] = p_cacc()[1] = p_cacc()[2] = 0.f;
      p_cacc()[3] = -P.opt[DO_gx]; p_cacc()[4] = -P.opt[DO_gy]; p_cacc()[5] = -P.opt[DO_gz];
    }
    sync();
    int maxdepth = dim(DD_maxdepth);
    for (int l = 1; l <= maxdepth; l++) {
      for (int idx = ladr[l] + lane; idx < ladr[l + 1]; idx += 32) {
        int b = lbody[idx], p = parent[b], ja = jadr[b], jn = jnum[b];
        V3 pos; Q4 quat;
        if (jn == 1 && jtype[ja] == 0) {
          float* q = p_qpos() + jq[ja];
          pos = ld3(q); quat = qnormalize(ldq(q + 3)); stq(q + 3, quat);
          int d = jd[ja];
          float mat[9]; quat2mat(mat, quat);
#pragma unroll
          for (int a = 0; a < 3; a++) {
            V3 e = v3(a == 0, a == 1, a == 2);
            st3(p_cdof() + 6 * (d + a), v3(0, 0, 0)); st3(p_cdof() + 6 * (d + a) + 3, e);
            V3 ax = matcol(mat, a);
            st3(p_cdof() + 6 * (d + 3 + a), ax); st3(p_cdof() + 6 * (d + 3 + a) + 3, cross(ax, pos) * -1.f);
          }
        } else {
          pos = ld3(p_xpos() + 3 * p) + mulmat(p_xmat() + 9 * p, ld3(bpos + 3 * b));
          quat = qmul(ldq(p_xquat() + 4 * p), ldq(bquat + 4 * b));
          for (int k = 0; k < jn; k++) {
            int j = ja + k, d = jd[j];
            V3 lp = ld3(jpos + 3 * j), la = ld3(jaxis + 3 * j);
            V3 anchor = qrot(quat, lp) + pos, axis = qrot(quat, la);
            float dq = p_qpos()[jq[j]] - q0[jq[j]];
            if (jtype[j] == 2) {  // slide
              pos = pos + axis * dq;
              st3(p_cdof() + 6 * d, v3(0, 0, 0)); st3(p_cdof() + 6 * d + 3, axis);
            } else {              // hinge
              quat = qmul(quat, axisangle(la, dq));
              pos = anchor - qrot(quat, lp);
              st3(p_cdof() + 6 * d, axis); st3(p_cdof() + 6 * d + 3, cross(axis, anchor) * -1.f);  // + axis x com later
            }
          }
          quat = qnormalize(quat);
        }
        st3(p_xpos() + 3 * b, pos); stq(p_xquat() + 4 * b, quat);
        float mat[9]; quat2mat(mat, quat);
#pragma unroll
        for (int k = 0; k < 9; k++) p_xmat()[9 * b + k] = mat[k];
        st3(p_xipos() + 3 * b, pos + mulmat(mat, ld3(bipos + 3 * b)));
      }
      sync();
    }
  }

  // ---- mj_comPos: com of each root subtree, cinert about it, cdof offset fix-up
  __device__ void com_pos() {
    const int* radr = I(DI_root_bodyadr); const int* rnum = I(DI_root_bodynum);
    const float* mass = F(DF_body_mass); const float* rinv = F(DF_root_invmass);
    int nroot = dim(DD_nroot);
    for (int r = lane; r < nroot; r += 32) {
      V3 acc = v3(0, 0, 0);
      int b0 = radr[r], n = rnum[r];
      for (int b = b0; b < b0 + n; b++) acc = acc + ld3(p_xipos() + 3 * b) * mass[b];
      if (rinv[r] > 0.f) acc = acc * rinv[r]; else acc = ld3(p_xipos() + 3 * b0);
      st3(p_rootcom() + 3 * r, acc);
    }
    sync();
    const int* ridx = I(DI_body_rootidx); const float* imat = F(DF_body_imat); const float* inertia = F(DF_body_inertia);
    int nb = dim(DD_nbody), nv = dim(DD_nv);
    for (int b = 1 + lane; b < nb; b += 32) {
      V3 dif = ld3(p_xipos() + 3 * b) - ld3(p_rootcom() + 3 * ridx[b]);
      const float* xm = p_xmat() + 9 * b; const float* im = imat + 9 * b;
      float R[9];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) R[3 * r + c] = xm[3 * r] * im[c] + xm[3 * r + 1] * im[3 + c] + xm[3 * r + 2] * im[6 + c];
      float i0 = inertia[3 * b], i1 = inertia[3 * b + 1], i2 = inertia[3 * b + 2], m = mass[b];
      float Ixx = R[0] * R[0] * i0 + R[1] * R[1] * i1 + R[2] * R[2] * i2;
      float Iyy = R[3] * R[3] * i0 + R[4] * R[4] * i1 + R[5] * R[5] * i2;
      float Izz = R[6] * R[6] * i0 + R[7] * R[7] * i1 + R[8] * R[8] * i2;
      float Ixy = R[0] * R[3] * i0 + R[1] * R[4] * i1 + R[2] * R[5] * i2;
      float Ixz = R[0] * R[6] * i0 + R[1] * R[7] * i1 + R[2] * R[8] * i2;
      float Iyz = R[3] * R[6] * i0 + R[4] * R[7] * i1 + R[5] * R[8] * i2;
      float d2 = dot(dif, dif);
      float* ci = p_cinert() + 10 * b;
      ci[0] = Ixx + m * (d2 - dif.x * dif.x); ci[1] = Iyy + m * (d2 - dif.y * dif.y); ci[2] = Izz + m * (d2 - dif.z * dif.z);
      ci[3] = Ixy - m * dif.x * dif.y; ci[4] = Ixz - m * dif.x * dif.z; ci[5] = Iyz - m * dif.y * dif.z;
      ci[6] = m * dif.x; ci[7] = m * dif.y; ci[8] = m * dif.z; ci[9] = m;
    }
    const int* isrot = I(DI_dof_isrot); const int* dbody = I(DI_dof_bodyid);
    for (int d = lane; d < nv; d += 32) {
      if (isrot[d]) {
        V3 ax = ld3(p_cdof() + 6 * d), com = ld3(p_rootcom() + 3 * ridx[dbody[d]]);
        st3(p_cdof() + 6 * d + 3, ld3(p_cdof() + 6 * d + 3) + cross(ax, com));
      }
    }
    sync();
  }

  // ---- forward velocity / bias-acceleration pass (mj_comVel + first half of mj_rne); cacc becomes the local cfrc
  __device__ void vel_pass() {
    const int* parent = I(DI_body_parentid); const int* jadr = I(DI_body_jntadr); const int* jnum = I(DI_body_jntnum);
    const int* jtype = I(DI_jnt_type); const int* jd = I(DI_jnt_dofadr);
    const int* ladr = I(DI_level_adr); const int* lbody = I(DI_level_body);
    int maxdepth = dim(DD_maxdepth), nb = dim(DD_nbody);
    for (int l = 1; l <= maxdepth; l++) {
      for (int idx = ladr[l] + lane; idx < ladr[l + 1]; idx += 32) {
        int b = lbody[idx], p = parent[b], ja = jadr[b], jn = jnum[b];
        S6 cv = ld6(p_cvel() + 6 * p), ca = ld6(p_cacc() + 6 * p);
        for (int k = 0; k < jn; k++) {
          int j = ja + k, d = jd[j];
          if (jtype[j] == 0) {
#pragma unroll
            for (int a = 0; a < 3; a++) cv = cv + ld6(p_cdof() + 6 * (d + a)) * p_qvel()[d + a];
            S6 cv0 = cv;
#pragma unroll
            for (int a = 3; a < 6; a++) {
              S6 cd = ld6(p_cdof() + 6 * (d + a)); float qv = p_qvel()[d + a];
              ca = ca + cross_motion(cv0, cd) * qv; cv = cv + cd * qv;
            }
          } else {
            S6 cd = ld6(p_cdof() + 6 * d); float qv = p_qvel()[d];
            ca = ca + cross_motion(cv, cd) * qv; cv = cv + cd * qv;
          }
        }
        st6(p_cvel() + 6 * b, cv); st6(p_cacc() + 6 * b, ca);
      }
      sync();
    }
    // cfrc_body = I*cacc + cvel x* (I*cvel), in place of cacc (children no longer need the parent's cacc)
    for (int b = lane; b < nb; b += 32) {
      S6 f; f.a = v3(0, 0, 0); f.l = v3(0, 0, 0);
      if (b > 0) {
        const float* ci = p_cinert() + 10 * b; S6 cv = ld6(p_cvel() + 6 * b), ca = ld6(p_cacc() + 6 * b);
        f = mul_inert(ci, ca) + cross_force(cv, mul_inert(ci, cv));
      }
      st6(p_cacc() + 6 * b, f);
    }
    sync();
  }

  // ---- fused backward pass: composite inertias in place of cinert (mj_crb) and subtree bias forces (mj_rne, part 2)
  __device__ void backward_pass() {
    const int* ladr = I(DI_level_adr); const int* lbody = I(DI_level_body);
    const int* cadr = I(DI_body_childadr); const int* cnum = I(DI_body_childnum); const int* child = I(DI_body_child);
    int maxdepth = dim(DD_maxdepth);
    for (int l = maxdepth - 1; l >= 1; l--) {
      for (int idx = ladr[l] + lane; idx < ladr[l + 1]; idx += 32) {
        int b = lbody[idx], n = cnum[b];
        if (!n) continue;
        float* crb = p_cinert() + 10 * b; float* cf = p_cacc() + 6 * b;
        for (int c = 0; c < n; c++) {
          int cb = child[cadr[b] + c];
#pragma unroll
          for (int k = 0; k < 10; k++) crb[k] += p_cinert()[10 * cb + k];
#pragma unroll
          for (int k = 0; k < 6; k++) cf[k] += p_cacc()[6 * cb + k];
        }
      }
      sync();
    }
  }

  // ---- sparse M, qfrc_smooth (passive - bias + applied + actuator)
  __device__ void mass_and_smooth() {
    const int* dbody = I(DI_dof_bodyid); const int* madr = I(DI_dof_Madr); const int* ddepth = I(DI_dof_depth);
    const int* mcol = I(DI_Mcol); const int* djnt = I(DI_dof_jnt); const int* jtype = I(DI_jnt_type);
    const int* jq = I(DI_jnt_qposadr); const int* aadr = I(DI_dof_actadr); const int* anum = I(DI_dof_actnum);
    const int* dofact = I(DI_dofact); const int* climited = I(DI_act_ctrllimited); const int* flimited = I(DI_act_forcelimited);
    const float* arm = F(DF_dof_armature); const float* damp = F(DF_dof_damping); const float* stiff = F(DF_jnt_stiffness);
    const float* qspring = F(DF_qpos_spring); const float* gear = F(DF_act_gear); const float* crange = F(DF_act_ctrlrange);
    const float* frange = F(DF_act_forcerange); const float* gain = F(DF_act_gain); const float* bias = F(DF_act_bias);
    int nv = dim(DD_nv);
    for (int i = lane; i < nv; i += 32) {
      S6 cd = ld6(p_cdof() + 6 * i);
      S6 buf = mul_inert(p_cinert() + 10 * dbody[i], cd);
      int a = madr[i], n = ddepth[i];
      p_M()[a] = dot6(cd, buf) + arm[i];
      for (int k = 1; k <= n; k++) p_M()[a + k] = dot6(ld6(p_cdof() + 6 * mcol[a + k]), buf);
      float qb = dot6(cd, ld6(p_cacc() + 6 * dbody[i]));
      float f = -damp[i] * p_qvel()[i] - qb + p_qapp()[i];
      int j = djnt[i];
      if (jtype[j] >= 2) { float k = stiff[j]; if (k != 0.f) f -= k * (p_qpos()[jq[j]] - qspring[jq[j]]); }
      for (int u = 0; u < anum[i]; u++) {
        int ac = dofact[aadr[i] + u];
        float c = p_ctrl()[ac];
        if (climited[ac]) c = clampf(c, crange[2 * ac], crange[2 * ac + 1]);
        float g = gear[ac];
        float af = gain[ac] * c + bias[3 * ac] + bias[3 * ac + 1] * g * p_qpos()[jq[j]] + bias[3 * ac + 2] * g * p_qvel()[i];
        if (flimited[ac]) af = clampf(af, frange[2 * ac], frange[2 * ac + 1]);
        f += g * af;
      }
      if (B.xfrc_body > 0) {        // mj_xfrcAccumulate for the one body the task pushes: [force(3), torque(3)] at its com
        int xb = B.xfrc_body;
        if ((I(DI_body_chainmask)[xb * dim(DD_nmaskw) + (i >> 5)] >> (i & 31)) & 1) {
          V3 fr = ld3(p_xfrc()), tq = ld3(p_xfrc() + 3);
          V3 off = ld3(p_xipos() + 3 * xb) - ld3(p_rootcom() + 3 * I(DI_body_rootidx)[xb]);
          f += dot(cd.a, tq + cross(off, fr)) + dot(cd.l, fr);
        }
      }
      p_qfs()[i] = f;
    }
    sync();
  }

  // ---- sparse L'DL factorisation of (M + hdamp*diag(damping)) (mj_factorM) as a flat op program:
  // one step per eliminated dof (leaves first); its rank-1 update is a list of independent (tgt, a, b) address
  // triples, one per lane; the division by D is deferred to one parallel pass, so a step costs one warp barrier and
  // three shared-memory round trips.
  __device__ __forceinline__ void factor(float hdamp, bool second) {
    const int* madr = I(DI_dof_Madr); const int* ddepth = I(DI_dof_depth);
    const int* fstep = I(DI_fac_step); const int* fops = I(DI_fac_ops); const float* damp = F(DF_dof_damping);
    float* LD = second ? p_LD2() : p_LD(); const float* M = p_M(); float* invD = second ? p_invD2() : p_invD();
    int nM = dim(DD_nM), nv = dim(DD_nv), nstep = dim(DD_nfacstep);
#pragma unroll 1
    for (int k = lane; k < nM; k += 32) LD[k] = M[k];
    sync();
    if (hdamp != 0.f) {
#pragma unroll 1
      for (int i = lane; i < nv; i += 32) LD[madr[i]] += hdamp * damp[i];
      sync();
    }
#pragma unroll 1
    for (int st = 0; st < nstep; st++) {
      int w = fstep[st], cnt = (w >> 10) & 255, oa = w >> 18;
      float inv = 1.0f / LD[w & 1023];
#pragma unroll 1
      for (int p = lane; p < cnt; p += 32) {
        int op = fops[oa + p];
        LD[op & 1023] -= LD[(op >> 10) & 1023] * inv * LD[op >> 20];
      }
      sync();
    }
#pragma unroll 1
    for (int i = lane; i < nv; i += 32) invD[i] = 1.0f / LD[madr[i]];
    sync();
    // deferred scaling: L(k, a_m) = LD(k, a_m) / D_k
#pragma unroll 1
    for (int i = lane; i < nv; i += 32) {
      int a = madr[i], n = ddepth[i]; float inv = invD[i];
      for (int m = 1; m <= n; m++) LD[a + m] *= inv;
    }
    sync();
  }

  // ---- x <- M^-1 x with the factor above; level-synchronous over dof depth (mj_solveLD).
  // backward (L^-T): every dof gathers from its descendants, deepest level first; forward (D^-1 then L^-1): a dof of
  // level l has exactly l ancestors, so the trip count is uniform across the level.
  __device__ __forceinline__ void solve(float* x, bool second) {
    const int* mcol = I(DI_Mcol); const int* dladr = I(DI_dlevel_adr);
    const int* bwp = I(DI_bw_pack); const int* fwp = I(DI_fw_pack); const int* dpack = I(DI_desc_pack);
    const float* LD = second ? p_LD2() : p_LD(); const float* invD = second ? p_invD2() : p_invD();
    int maxd = dim(DD_maxdofdepth);
#pragma unroll 1
    for (int l = maxd - 1; l >= 0; l--) {
#pragma unroll 1
      for (int idx = dladr[l] + lane; idx < dladr[l + 1]; idx += 32) {
        int w = bwp[idx], j = w & 255, n = (w >> 8) & 255; const int* dp = dpack + (w >> 16);
        float s0 = x[j], s1 = 0.f, s2 = 0.f, s3 = 0.f; int k = 0;
#pragma unroll 1
        for (; k + 4 <= n; k += 4) {
          int p0 = dp[k], p1 = dp[k + 1], p2 = dp[k + 2], p3 = dp[k + 3];
          s0 = fmaf(-LD[p0 >> 16], x[p0 & 0xffff], s0); s1 = fmaf(-LD[p1 >> 16], x[p1 & 0xffff], s1);
          s2 = fmaf(-LD[p2 >> 16], x[p2 & 0xffff], s2); s3 = fmaf(-LD[p3 >> 16], x[p3 & 0xffff], s3);
        }
#pragma unroll 1
        for (; k < n; k++) { int p0 = dp[k]; s0 = fmaf(-LD[p0 >> 16], x[p0 & 0xffff], s0); }
        x[j] = (s0 + s1) + (s2 + s3);
      }
      sync();
    }
#pragma unroll 1
    for (int l = 0; l <= maxd; l++) {
#pragma unroll 1
      for (int idx = dladr[l] + lane; idx < dladr[l + 1]; idx += 32) {
        int w = fwp[idx], i = w & 255, a = w >> 8;
        float s0 = x[i] * invD[i], s1 = 0.f; int k = 1;
#pragma unroll 1
        for (; k + 1 <= l; k += 2) { s0 = fmaf(-LD[a + k], x[mcol[a + k]], s0); s1 = fmaf(-LD[a + k + 1], x[mcol[a + k + 1]], s1); }
        if (k <= l) s0 = fmaf(-LD[a + k], x[mcol[a + k]], s0);
        x[i] = s0 + s1;
      }
      sync();
    }
  }

  // ---- B.4 narrow phase, one candidate pair per lane, raw contacts into the arena, ordered compaction
  __device__ __forceinline__ void geom_pose(int cg, V3& pos, float* mat) const {
    const int* cgbody = I(DI_cg_body); const float* cgpos = F(DF_cg_pos); const float* cgmat = F(DF_cg_mat);
    int b = cgbody[cg];
    const float* xm = p_xmat() + 9 * b; const float* lm = cgmat + 9 * cg;
    pos = ld3(p_xpos() + 3 * b) + mulmat(xm, ld3(cgpos + 3 * cg));
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) mat[3 * r + c] = xm[3 * r] * lm[c] + xm[3 * r + 1] * lm[3 + c] + xm[3 * r + 2] * lm[6 + c];
  }
  // true when the point c is farther than `reach` from the box geom (cg, on body b, centre cb)
  __device__ __forceinline__ bool box_far(int cg, int b, V3 cb, V3 c, float reach) const {
    const float* xm = p_xmat() + 9 * b; const float* lm = F(DF_cg_mat) + 9 * cg; const float* sz = F(DF_cg_size) + 3 * cg;
    V3 dw = c - cb; V3 db = mulmatT(xm, dw); V3 dl = mulmatT(lm, db);      // into the body frame, then the geom frame
    float ex = fmaxf(fabsf(dl.x) - sz[0], 0.f), ey = fmaxf(fabsf(dl.y) - sz[1], 0.f), ez = fmaxf(fabsf(dl.z) - sz[2], 0.f);
    return ex * ex + ey * ey + ez * ez > reach * reach;
  }
  // candidate pairs are culled first (bounding spheres / plane distance) and the survivors compacted, in pair order,
  // into an active list, so the narrow phase runs one *surviving* pair per lane.  Raw contact slots are handed out to the
  // active pairs by a prefix sum of their static maxima, so the arena only holds what a step can actually produce
  // (capacities act_cap / raw_cap; what does not fit is counted as dropped).
  __device__ void collision(unsigned long long* counters) {
    const int* pc1 = PI(DI_pair_cg1); const int* pc2 = PI(DI_pair_cg2); const int* pprm = PI(DI_pair_prm);
    const int* pmax = PI(DI_pair_maxcon); const int* cgtype = I(DI_cg_type);
    const int* cgbody = I(DI_cg_body); const float* cgpos = F(DF_cg_pos); const float* cgmat = F(DF_cg_mat);
    const float* cgsize = F(DF_cg_size); const float* cgrb = F(DF_cg_rbound); const float* prm = F(DF_prm);
    const int npair = dim(DD_npair), acap = B.act_cap, rcap = B.raw_cap;
    // arena: [active pair ids | raw slot of each | contact count of each | 10 floats per raw contact slot]
    int* alist = (int*)p_arena(); int* araw = alist + acap; int* rcount = araw + acap; float* rdata = p_arena() + 3 * acap;
    const unsigned lt = (1u << lane) - 1u;
    int nact = 0, dropped = 0;
#ifdef B2_PHASE_TIMING
    long long tc_ = clock64();      // [4] cull + slot hand-out, [5] narrow phase; the rest of [8] is the ordered compaction
#define B2_CT(slot) do { if (lane == 0 && B.phase_cycles) { const long long t_ = clock64(); atomicAdd(&B.phase_cycles[slot], (unsigned long long)(t_ - tc_)); tc_ = t_; } } while (0)
#else
#define B2_CT(slot) do { } while (0)
#endif
    for (int p0 = 0; p0 < npair; p0 += 32) {
      int p = p0 + lane; bool keep = false;
      if (p < npair) {
        int g1 = pc1[p], g2 = pc2[p]; float margin = prm[B2DEV_PRM_STRIDE * pprm[p]];
        int b1 = cgbody[g1], b2 = cgbody[g2];
        V3 q1 = ld3(p_xpos() + 3 * b1) + mulmat(p_xmat() + 9 * b1, ld3(cgpos + 3 * g1));
        V3 q2 = ld3(p_xpos() + 3 * b2) + mulmat(p_xmat() + 9 * b2, ld3(cgpos + 3 * g2));
        V3 d = q2 - q1;
        if (cgtype[g1] == 0) {
          V3 n = mulmat(p_xmat() + 9 * b1, matcol(cgmat + 9 * g1, 2));
          keep = !(dot(d, n) > cgrb[g2] + margin);
        } else {
          float bd = cgrb[g1] + cgrb[g2] + margin; keep = !(dot(d, d) > bd * bd);
          // a long box has a huge bounding sphere: also test the other geom's bounding sphere against the box itself
          // (pure pruning: the narrow phase would return nothing for these pairs)
          if (keep && cgtype[g2] == GT_BOX) keep = !box_far(g2, b2, q2, q1, cgrb[g1] + margin);
          if (keep && cgtype[g1] == GT_BOX) keep = !box_far(g1, b1, q1, q2, cgrb[g2] + margin);
        }
      }
      unsigned m = __ballot_sync(B2_FULL, keep);
      int slot = nact + __popc(m & lt);
      if (keep) { if (slot < acap) alist[slot] = p; else dropped += pmax[p]; }
      nact = min(nact + __popc(m), acap);
    }
    sync();
    // raw slots: exclusive prefix sum of the active pairs' maxima; a pair whose slots do not all fit gets none
    int rbase = 0;
    for (int k0 = 0; k0 < nact; k0 += 32) {
      int k = k0 + lane; int n = (k < nact) ? pmax[alist[k]] : 0;
      int incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(B2_FULL, incl, o); if (lane >= o) incl += v; }
      int start = rbase + incl - n;
      if (k < nact) { if (start + n <= rcap) araw[k] = start; else { araw[k] = -1; dropped += n; } }
      rbase += __shfl_sync(B2_FULL, incl, 31);
    }
    sync();
    B2_CT(4);
    for (int k = lane; k < nact; k += 32) {
      int p = alist[k], g1 = pc1[p], g2 = pc2[p]; float margin = prm[B2DEV_PRM_STRIDE * pprm[p]];
      int n = 0;
      if (araw[k] >= 0) {
        V3 p1, p2; float m1[9], m2[9];
        geom_pose(g1, p1, m1); geom_pose(g2, p2, m2);
        n = collide_pair<CVX>(cgtype[g1], cgtype[g2], p1, m1, cgsize + 3 * g1, p2, m2, cgsize + 3 * g2, margin,
                         rdata + B2_RAW * araw[k], pmax[p]);
      }
      rcount[k] = n;
    }
    sync();
    B2_CT(5);
    // tier of this forward pass: more contacts than the on-chip buffer holds -> wide workspace (nothing is dropped)
    {
      int tot = 0;
      for (int k = lane; k < nact; k += 32) tot += rcount[k];
#pragma unroll
      for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(B2_FULL, tot, o);
      wide = B.wide != nullptr && tot > B.con_cap;
      if (lane == 0) { p_misc()[MISC_WIDE] = wide ? 1 : 0; if (wide && counters) atomicAdd(&counters[CTR_WIDE], 1ull); }
    }
    float* const conbuf = x_con();
    // ordered compaction (active list is in pair order == MuJoCo contact order)
    int base = 0;
    for (int k0 = 0; k0 < nact; k0 += 32) {
      int k = k0 + lane; int n = (k < nact) ? rcount[k] : 0; int p = (k < nact) ? alist[k] : 0;
      int incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(B2_FULL, incl, o); if (lane >= o) incl += v; }
      int start = base + incl - n;
      for (int q = 0; q < n; q++) {
        int c = start + q;
        if (c >= conCap()) { dropped++; continue; }
        const float* src = rdata + B2_RAW * (araw[k] + q);
        float* dst = conbuf + B2_CON_STRIDE * c;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        // frame: normal, then orthogonalised tangent hint (mju_makeFrame)
        V3 nrm = normalized(ld3(src + 4)); V3 t = ld3(src + 7);
        if (norm(t) < 0.5f) { t = (nrm.y < 0.5f && nrm.y > -0.5f) ? v3(0, 1, 0) : v3(0, 0, 1); }
        t = t - nrm * dot(nrm, t); t = normalized(t);
        V3 t2 = cross(nrm, t);
        st3(dst + 4, nrm); st3(dst + 7, t); st3(dst + 10, t2);
        dst[13] = __int_as_float(p); dst[14] = 0.f; dst[15] = 0.f;
      }
      base += __shfl_sync(B2_FULL, incl, 31);
    }
    dropped = (int)warp_sum((float)dropped);
    if (lane == 0) {
      p_misc()[MISC_NCON] = base < conCap() ? base : conCap();
      if (dropped && counters) atomicAdd(&counters[CTR_CON_DROPPED], (unsigned long long)dropped);
    }
    sync();
  }

  // island of contact c: the island of its moving body (geom 2's tree when it has one, else geom 1's)
  __device__ __forceinline__ int contact_island(int c) const {
    const int* btree = I(DI_body_tree); const int* cgbody = I(DI_cg_body);
    int p = __float_as_int(x_con()[B2_CON_STRIDE * c + 13]);
    int t2 = btree[cgbody[PI(DI_pair_cg2)[p]]], t1 = btree[cgbody[PI(DI_pair_cg1)[p]]];
    return p_tree_isl()[t2 >= 0 ? t2 : t1];
  }
  // column -> dof of island k.  Islands made of one tree (or of adjacent trees) cover a contiguous dof range, which
  // is the common case; then the map is an add instead of a shared-memory load.
  struct Cols {
    const int* map; int d0; bool contig;
    __device__ __forceinline__ int dof(int c) const { return (!DYN || contig) ? d0 + c : map[c]; }
  };
  __device__ __forceinline__ Cols island_cols(int k) const {
    Cols q; q.map = p_col_dof() + p_isl_col0()[k];
    if (!DYN) { q.d0 = p_isl_col0()[k]; q.contig = true; return q; }       // static islands are contiguous dof ranges
    int nd = p_isl_nd()[k];
    q.d0 = q.map[0]; q.contig = q.map[nd - 1] - q.d0 == nd - 1;
    return q;
  }
  // static islands (models whose moving trees never share a candidate pair): set up once per launch
  __device__ void static_islands() {
    const int* disl = I(DI_dof_island); const int* iadr = I(DI_island_dofadr); const int* inum = I(DI_island_dofnum);
    const int* tadr = I(DI_tree_dofadr);
    int ntree = dim(DD_ntree), nv = dim(DD_nv), nisl = dim(DD_nisland);
    if (lane < ntree) { p_tree_isl()[lane] = disl[tadr[lane]]; p_tree_col0()[lane] = tadr[lane] - iadr[disl[tadr[lane]]]; }
    if (lane < nisl) { p_isl_nd()[lane] = inum[lane]; p_isl_col0()[lane] = iadr[lane]; }
    if (lane == 0) { p_isl_col0()[nisl] = nv; p_misc()[MISC_NISL] = nisl; }
    for (int d = lane; d < nv; d += 32) { p_dof_col()[d] = d - iadr[disl[d]]; p_col_dof()[d] = d; }
    sync();
  }
  // ---- islands of this forward pass: kinematic trees joined by a contact between two moving bodies (union-find over at
  // most 16 trees, by one lane), numbered by their smallest tree; an island's columns are its trees' dofs in tree order.
  // Islands are exactly decoupled blocks of the constraint problem, so solving them separately, each in MuJoCo's row
  // order, is the same computation as MuJoCo's single sweep over all rows.
  __device__ void make_islands() {
    const int* btree = I(DI_body_tree); const int* cgbody = I(DI_cg_body); const int* pc1 = PI(DI_pair_cg1); const int* pc2 = PI(DI_pair_cg2);
    const int* tadr = I(DI_tree_dofadr); const int* tnum = I(DI_tree_dofnum); const int* dtree = I(DI_dof_tree);
    if (!DYN) return;
    int ntree = dim(DD_ntree), nv = dim(DD_nv), ncon = p_misc()[MISC_NCON];
    int* lab = p_tree_isl();
    if (ntree == 1) {                              // one tree: one island, columns = dofs
      if (lane == 0) { lab[0] = 0; p_tree_col0()[0] = 0; p_isl_nd()[0] = nv; p_isl_col0()[0] = 0; p_isl_col0()[1] = nv; p_misc()[MISC_NISL] = 1; }
      if (p_col_dof()[nv - 1] != nv - 1 || p_dof_col()[nv - 1] != nv - 1) { for (int d = lane; d < nv; d += 32) { p_dof_col()[d] = d; p_col_dof()[d] = d; } }
      sync();
      return;
    }
    if (lane < ntree) lab[lane] = lane;
    sync();
    // lanes look their contacts' trees up in parallel; only contacts between two moving bodies reach the serial union
    for (int c0 = 0; c0 < ncon; c0 += 32) {
      int c = c0 + lane; int a = -1, b = -1;
      if (c < ncon) { int p = __float_as_int(x_con()[B2_CON_STRIDE * c + 13]); a = btree[cgbody[pc1[p]]]; b = btree[cgbody[pc2[p]]]; }
      unsigned m = __ballot_sync(B2_FULL, a >= 0 && b >= 0 && a != b);
      while (m) {
        int src = __ffs(m) - 1; m &= m - 1;
        int ua = __shfl_sync(B2_FULL, a, src), ub = __shfl_sync(B2_FULL, b, src);
        if (lane == 0) {
          while (lab[ua] != ua) ua = lab[ua];
          while (lab[ub] != ub) ub = lab[ub];
          if (ua != ub) lab[max(ua, ub)] = min(ua, ub);
        }
      }
    }
    sync();
    if (lane == 0) {
      int nisl = 0;
      for (int t = 0; t < ntree; t++) {            // roots are the smallest tree of their set: number islands in tree order
        int r = t; while (lab[r] != r) r = lab[r];
        if (r == t) { p_isl_warp()[t] = nisl; p_isl_nd()[nisl] = 0; nisl++; }     // isl_warp doubles as root -> island id here
        p_tree_col0()[t] = r;                                                       // remember the root
      }
      for (int t = 0; t < ntree; t++) {
        int k = p_isl_warp()[p_tree_col0()[t]];
        lab[t] = -1 - k;                           // final ids are written below (kept negative until every root is read)
        p_tree_col0()[t] = p_isl_nd()[k]; p_isl_nd()[k] += tnum[t];
      }
      int col0 = 0;
      for (int k = 0; k < nisl; k++) { p_isl_col0()[k] = col0; col0 += p_isl_nd()[k]; }
      p_isl_col0()[nisl] = col0;
      for (int t = 0; t < ntree; t++) lab[t] = -1 - lab[t];
      p_misc()[MISC_NISL] = nisl;
    }
    sync();
    for (int d = lane; d < nv; d += 32) {
      int t = dtree[d], k = lab[t], c = p_tree_col0()[t] + (d - tadr[t]);
      p_dof_col()[d] = c; p_col_dof()[p_isl_col0()[k] + c] = d;
    }
    sync();
  }

  // rows of contact c: 2 (condim - 1)
  __device__ __forceinline__ int contact_rows(int c) const {
    if (!C6) return 4;
    int p = __float_as_int(x_con()[B2_CON_STRIDE * c + 13]);
    return F(DF_prm)[B2DEV_PRM_STRIDE * PI(DI_pair_prm)[p] + 14] == 6.0f ? 10 : 4;
  }
  // ---- B.5 rows: joint limits then pyramidal contacts, stably partitioned by island.
  // Rows are numbered without a cap first; the carve-up then picks the tier of this forward pass:
  //   fast  every island's J and packed A (Newton: J, H and its vectors) fit the on-chip arena, <= 128 rows per PGS island;
  //   wide  contact records, row arrays and J live in the env's global workspace (B.wide); PGS sweeps matrix-free over
  //         [J | M^-1 J' | record] blocks streamed through a cp.async ring, Newton keeps H on chip and reads J from global.
  // Only what exceeds the wide capacities as well is cut from an island's tail (whole pyramids, last contacts first) and counted.
  __device__ __forceinline__ int wide_blkf_max() const { return 8 * r4(dim(DD_maxspan)) + B2_WREC; }
  __device__ void make_rows(unsigned long long* counters) {
    const int* limj = I(DI_lim_jnt); const int* jq = I(DI_jnt_qposadr); const int* jd = I(DI_jnt_dofadr);
    const int* dtree = I(DI_dof_tree);
    const float* jrange = F(DF_jnt_range); const float* jmargin = F(DF_jnt_margin);
    make_islands();
    int nlim = dim(DD_nlim), nisl = p_misc()[MISC_NISL], ncon = p_misc()[MISC_NCON];
    if (lane < B2_MAX_ISLANDS) p_isl_n()[lane] = 0;
    sync();
    unsigned lt = (1u << lane) - 1u;
    for (int c0 = 0; c0 < nlim; c0 += 32) {
      int k = c0 + lane; bool valid = k < nlim;
      int j = valid ? limj[k] : 0;
      float q = p_qpos()[jq[j]], mg = jmargin[j];
      bool lo = valid && (q - jrange[2 * j] < mg), hi = valid && (jrange[2 * j + 1] - q < mg);
      int isl = valid ? p_tree_isl()[dtree[jd[j]]] : -1 - lane;
      unsigned peers = __match_any_sync(B2_FULL, isl);
      unsigned lomask = __ballot_sync(B2_FULL, lo), himask = __ballot_sync(B2_FULL, hi);
      int before = __popc(lomask & peers & lt) + __popc(himask & peers & lt);
      int total = __popc(lomask & peers) + __popc(himask & peers);
      int base = valid ? p_isl_n()[isl] : 0;
      sync();
      int r0 = base + before;
      if (valid) {
        int rl = -1, rh = -1;
        if (lo) { rl = r0; r0++; }
        if (hi) rh = r0;
        p_lim_row()[2 * k] = rl; p_lim_row()[2 * k + 1] = rh;
        if ((peers & lt) == 0 && total) p_isl_n()[isl] = base + total;
      }
      sync();
    }
    if (lane < B2_MAX_ISLANDS) p_isl_nl()[lane] = p_isl_n()[lane];      // limit rows come first in every island
    sync();
    int* const conrow = x_con_row();
    for (int c0 = 0; c0 < ncon; c0 += 32) {
      int c = c0 + lane; bool valid = c < ncon;
      int isl = valid ? contact_island(c) : -1 - lane;
      unsigned peers = __match_any_sync(B2_FULL, isl);
      int before = 4 * __popc(peers & lt), total = 4 * __popc(peers);
      int base = valid ? p_isl_n()[isl] : 0;
      const int nr = valid ? contact_rows(c) : 0;
      if (C6) {        // pyramids of 4 or 10 rows: weighted prefix over the lanes of the same island
        before = 0; total = 0;
        for (int src = 0; src < 32; src++) { int w = __shfl_sync(B2_FULL, nr, src); if ((peers >> src) & 1) { total += w; if (src < lane) before += w; } }
      }
      sync();
      if (valid) {
        conrow[c] = base + before;
        if ((peers & lt) == 0) p_isl_n()[isl] = base + total;
      }
      sync();
    }
    // tier, island bases and arena carve-up
    if (lane == 0) {
      const bool nw = newton(); const int scratch = scratch_in_arena(); const int nv = dim(DD_nv);
      bool wd = wide;
      if (!wd && B.wide) {
        int adr = 0, used = 0; bool ok = true;
        for (int k = 0; k < nisl; k++) {
          int n = p_isl_n()[k], ndk = p_isl_nd()[k], ldj = ndk | 1;
          if (n > (nw ? B.row_cap : B2_ISLAND_ROWS)) ok = false;
          adr += n; used += r4(n * ldj) + (nw ? newton_floats(n, ndk) : a_floats(n));
        }
        if (adr > B.row_cap || used > arenaFloats() - scratch - 8) ok = false;
        if (!ok) { wd = true; if (counters) { atomicAdd(&counters[CTR_WIDE], 1ull); atomicAdd(&counters[CTR_WIDE_ROWS], 1ull); } }
      }
      const int rcap = wd ? B.w_row_cap : B.row_cap;
      const int gavail = wd ? B.w_arena_floats - 8 : arenaFloats() - scratch - 8;     // where J goes
      int savail = arenaFloats() - 8, sbase = 0, nscr = W;                            // on-chip part of the wide tier
      if (wd && !nw) {
        // on chip: scratch columns for the B build, then every island's f and v, then the bulk-copy rings of the sweeping warps
        // (at least RING_MIN stages each are reserved here; the depth actually used is whatever is left, below)
        const int ringf = W * RING_MIN * (wide_blkf_max() + 4);
        int fv = 0;
        for (int k = 0; k < nisl; k++) fv += r4(p_isl_n()[k]) + 4 + r4(p_isl_nd()[k]);
        if (nscr * 32 * nv + ringf + fv > savail) nscr = 1;
        sbase = nscr * 32 * nv; savail -= sbase + ringf;
      }
      int adr = 0, used = 0, sused = 0, ovf = 0, cut = 0;
      for (int k = 0; k < nisl; k++) {
        const int n0 = p_isl_n()[k], ndk = p_isl_nd()[k], nl = p_isl_nl()[k];
        const int maxrows = (wd || nw) ? rcap : B2_ISLAND_ROWS;       // the 128-row island limit is the fast PGS register layout's
        const int ldj = (wd && !nw) ? r4(ndk) : (ndk | 1);
        auto needJ = [&](int n) { return (wd && !nw) ? ((n + 3) >> 2) * (8 * ldj + B2_WREC) : (wd ? r4(n * ldj) + r4(n) : r4(n * ldj)); };
        auto needS = [&](int n) { return !wd ? (nw ? newton_floats(n, ndk) : a_floats(n)) : (nw ? newton_floats(0, ndk) : r4(n) + 4 + r4(ndk)); };
        auto fits = [&](int n) { return wd ? (used + needJ(n) <= gavail && sused + needS(n) <= savail) : (used + needJ(n) + needS(n) <= gavail); };
        int n = min(n0, maxrows);
        if (adr + n > rcap) n = max(rcap - adr, 0);
        while (n > 0 && !fits(n)) n--;
        if (n < n0) {
          // rows that do not fit are cut from the island's tail: the last contacts go first, the joint limits last; whole pyramids only
          ovf++;
          if (!C6) { if (n > nl) n = nl + ((n - nl) >> 2) * 4; }
          else if (n > nl) {
            int nb = nl;
            for (int c = 0; c < ncon; c++) {
              int r = conrow[c];
              if (contact_island(c) == k) { int e = r + contact_rows(c); if (e <= n && e > nb) nb = e; }
            }
            n = nb;
          }
          cut += n0 - n;
        }
        p_isl_n()[k] = n; p_isl_adr()[k] = adr; p_isl_ldj()[k] = ldj; p_isl_J()[k] = used;
        if (!wd) { p_isl_A()[k] = used + needJ(n); used += needJ(n) + needS(n); }
        else { p_isl_A()[k] = sbase + sused; used += needJ(n); sused += needS(n); }
        adr += n;
      }
      p_isl_adr()[nisl] = adr; p_misc()[MISC_NEFC] = adr; p_misc()[MISC_ARENA_USED] = used;
      p_misc()[MISC_WIDE] = wd ? 1 : 0; p_misc()[MISC_WSCR] = nscr;
      // islands -> warps of the team, greedily by sweep length (4-row blocks), so the A build and the PGS sweeps of
      // one env finish together
      int load[W > 1 ? W : 1];
#pragma unroll
      for (int q = 0; q < W; q++) load[q] = 0;
      for (int k = 0; k < nisl; k++) {
        int best = 0;
#pragma unroll
        for (int q = 1; q < W; q++) if (load[q] < load[best]) best = q;
        p_isl_warp()[k] = best;
#pragma unroll
        for (int q = 0; q < W; q++) if (q == best) load[q] += p_isl_n()[k] ? ((p_isl_n()[k] + 3) >> 2) + 2 : 0;
      }
      if (wd && !nw) {
        int nsw = 0;
#pragma unroll
        for (int q = 0; q < W; q++) nsw += load[q] > 0 ? 1 : 0;
        const int rb = sbase + sused, left = arenaFloats() - 8 - rb;
        int depth = nsw ? left / (nsw * (wide_blkf_max() + 4)) : RING_MIN;
        p_misc()[MISC_WRING] = max(RING_MIN, min(RING_MAX, depth)); p_misc()[MISC_WRB] = rb;
      }
      if (counters) {
        if (ovf) atomicAdd(&counters[CTR_ARENA_OVERFLOW], (unsigned long long)ovf);
        if (cut) atomicAdd(&counters[CTR_ROW_DROPPED], (unsigned long long)cut);
      }
    }
    sync();
    if (!wide && p_misc()[MISC_WIDE]) {      // upgraded by the carve-up: move the contact records and their row numbers over
      const float* sc = b2_smem + wb + B.off.con; const int* sr = (const int*)(b2_smem + wb + B.off.con_row);
      float* gc = wbase() + B.woff.con; int* gr = (int*)(wbase() + B.woff.con_row);
      for (int i = lane; i < ncon * B2_CON_STRIDE; i += 32) gc[i] = sc[i];
      for (int i = lane; i < ncon; i += 32) gr[i] = sr[i];
      wide = true;
      sync();
    }
  }
  // column scratch for the A build: 32 * max island dof span floats; lives in the dead block when it fits
  __device__ __forceinline__ int max_span() const { return dim(DD_maxspan); }
  // A-build scratch: 32 floats per dof; in the dead block [xmat .. cinert] when it fits, else at the arena tail
  __device__ __forceinline__ int scratch_in_arena() const {
    return (newton() || 32 * dim(DD_nv) <= dead_block_floats(P.dim, B.keep_frames)) ? 0 : 32 * dim(DD_nv);
  }
  __device__ __forceinline__ float* scratch_base() const {
    int sc = scratch_in_arena();
    return sc ? p_arena() + arenaFloats() - sc : p_xmat();
  }
  __device__ __forceinline__ float* island_A(int k) const { return p_arena() + p_isl_A()[k]; }

  // ---- fill J (island-dense), per-row parameters, aref, b, warm-start force
  template <bool WD> __device__ void fill_rows() {
    const int* limj = I(DI_lim_jnt); const int* jd = I(DI_jnt_dofadr); const int* jq = I(DI_jnt_qposadr);
    const int* dtree = I(DI_dof_tree); const int* cgbody = I(DI_cg_body);
    const int* pc1 = PI(DI_pair_cg1); const int* pc2 = PI(DI_pair_cg2); const int* pprm = PI(DI_pair_prm);
    const int* dbody = I(DI_dof_bodyid);
    const int* ridx = I(DI_body_rootidx); const int* cmask = I(DI_body_chainmask);
    const float* jrange = F(DF_jnt_range); const float* jmargin = F(DF_jnt_margin); const float* jsol = F(DF_jnt_solprm);
    const float* dinvw = F(DF_dof_invweight0); const float* binvw = F(DF_body_invweight0); const float* prm = F(DF_prm);
    int nlim = dim(DD_nlim), nisl = p_misc()[MISC_NISL], ncon = p_misc()[MISC_NCON], nmw = dim(DD_nmaskw);
    float timestep = P.opt[DO_timestep], impratio = P.opt[DO_impratio];
    int* const rinfo = xs_row_info<WD>(); int* const conrow = xs_con_row<WD>(); const float* const conbuf = xs_con<WD>();
    float* const rowR = xs_row_R<WD>(); float* const rowb = xs_row_b<WD>(); float* const rowf = xs_row_f<WD>(); float* const rowres = xs_row_res<WD>();
    // row_info: limits  -> (joint << 4) | side ; contacts -> 0x40000000 | (contact << 4) | row of the pyramid
    for (int k = tl; k < 2 * nlim; k += TEAM) {
      int r = p_lim_row()[k]; if (r < 0) continue;
      int j = limj[k >> 1], isl = p_tree_isl()[dtree[jd[j]]];
      if (r >= p_isl_n()[isl]) continue;
      rinfo[p_isl_adr()[isl] + r] = (j << 4) | (k & 1);
    }
    for (int c = tl; c < ncon; c += TEAM) {
      int r = conrow[c]; if (r < 0) continue;
      int isl = contact_island(c);
      const int nr = contact_rows(c);
      if (r + nr > p_isl_n()[isl]) { conrow[c] = -1; continue; }
      for (int d = 0; d < nr; d++) rinfo[p_isl_adr()[isl] + r + d] = 0x40000000 | (c << 4) | d;
    }
    if (WD && !newton()) {      // wide PGS blocks: padding columns, rows past n and the B / record parts start from zero
      for (int k = 0; k < nisl; k++) {
        int n = p_isl_n()[k]; if (!n) continue;
        float4* z = reinterpret_cast<float4*>(xs_J<WD>(k)); const int nz = ((n + 3) >> 2) * (jblk(p_isl_ldj()[k]) >> 2);
        for (int i = tl; i < nz; i += TEAM) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    team_sync();
    // One thread per row: the row's parameters (impedance, R, reference acceleration) and its J entries, with the products of
    // the row with qvel, qacc_smooth and qacc_warmstart accumulated while the entries are formed.  Everything that depends on
    // the row only (contact record, pair parameters, the direction the pyramid row pushes along, the lever arms to the two
    // bodies' tree roots) is loaded once per row; an entry then costs the chain-mask test, one 6-float cdof load, a cross
    // product and a dot product: J_ed = sgn (u . (cdof_lin + cdof_ang x off)), the same arithmetic as before.
    // (Round 2's first form took one (row, dof) entry per thread and re-walked the chain rinfo -> contact -> pair -> bodies ->
    // masks -> parameters for each: eight dependent shared-memory loads per entry, 30-40 % of the arm's sub-step.)
    // per-row parameters
    int nefc = p_misc()[MISC_NEFC];
    for (int e = tl; e < nefc; e += TEAM) {
      int info = rinfo[e];
      float pos, margin, da, solref0, solref1; const float* simp; int isl; float mu0 = 0.f; bool iscon = info & 0x40000000;
      if (iscon) {
        int ci = (info >> 4) & 0x03ffffff;
        const float* con = conbuf + B2_CON_STRIDE * ci; int p = __float_as_int(con[13]);
        const float* pr = prm + B2DEV_PRM_STRIDE * pprm[p];
        int b1 = cgbody[pc1[p]], b2 = cgbody[pc2[p]];
        isl = contact_island(ci);
        pos = con[0]; margin = pr[0] - pr[1]; mu0 = pr[2];
        float tran = binvw[2 * b1] + binvw[2 * b2];
        da = tran + mu0 * mu0 * tran;   // first row of the pyramid sets R for all of its rows
        solref0 = pr[7]; solref1 = pr[8]; simp = pr + 9;
      } else {
        int j = info >> 4, side = info & 1; float q = p_qpos()[jq[j]];
        isl = p_tree_isl()[dtree[jd[j]]];
        pos = side ? jrange[2 * j + 1] - q : q - jrange[2 * j]; margin = jmargin[j]; da = dinvw[jd[j]];
        solref0 = jsol[8 * j]; solref1 = jsol[8 * j + 1]; simp = jsol + 8 * j + 2;
      }
      da = fmaxf(da, B2_MINVAL);
      // impedance (getimpedance)
      float dmin = clampf(simp[0], B2_MINIMP, B2_MAXIMP), dmax = clampf(simp[1], B2_MINIMP, B2_MAXIMP);
      float width = fmaxf(simp[2], 0.f), mid = clampf(simp[3], B2_MINIMP, B2_MAXIMP), power = fmaxf(simp[4], 1.f);
      float imp;
      if (dmin == dmax || width <= B2_MINVAL) imp = 0.5f * (dmin + dmax);
      else {
        float x = fabsf(pos - margin) / width;
        if (x >= 1.f) imp = dmax;
        else if (x <= 0.f) imp = dmin;
        else {
          float y;
          if (power == 1.f) y = x;
          else if (power == 2.f) y = (x <= mid) ? x * x / mid : 1.f - (1.f - x) * (1.f - x) / (1.f - mid);
          else if (x <= mid) y = __powf(x, power) / __powf(mid, power - 1.f);
          else y = 1.f - __powf(1.f - x, power) / __powf(1.f - mid, power - 1.f);
          imp = dmin + y * (dmax - dmin);
        }
      }
      float R = fmaxf((1.f - imp) * da / imp, B2_MINVAL);
      if (iscon) { float mu = mu0 * sqrtf(1.f / impratio); R = fmaxf(2.f * mu * mu * R, B2_MINVAL); }
      float K, B;
      if (solref0 > 0.f) {
        float tc = fmaxf(solref0, 2.f * timestep);
        K = 1.f / fmaxf(dmax * dmax * tc * tc * solref1 * solref1, B2_MINVAL); B = 2.f / fmaxf(dmax * tc, B2_MINVAL);
      } else { K = -solref0 / (dmax * dmax); B = -solref1 / dmax; }
      // J row and its products with qvel, qacc_smooth, qacc_warmstart
      int nd = p_isl_nd()[isl], ldj = p_isl_ldj()[isl]; const Cols cols = island_cols(isl);
      float* Jr = xs_J<WD>(isl) + jr<WD>(e - p_isl_adr()[isl], jblk(ldj), ldj);
      float vel = 0.f, ja = 0.f, jw = 0.f;
      if (iscon) {
        const int ci = (info >> 4) & 0x03ffffff, dir = info & 15, kd = dir >> 1;      // kd 0, 1: translation along t1, t2; 2, 3, 4 (condim 6): rotation about n, t1, t2
        const float* con = conbuf + B2_CON_STRIDE * ci; const int p = __float_as_int(con[13]);
        const int b1 = cgbody[pc1[p]], b2 = cgbody[pc2[p]];
        float mu = (prm + B2DEV_PRM_STRIDE * pprm[p])[2 + kd]; mu = (dir & 1) ? -mu : mu;
        const V3 nrm = ld3(con + 4), cpos = ld3(con + 1);
        // direction the row pushes along (translation rows) / the axis it turns about (condim-6 rotation rows), and the lever
        // arms from the two bodies' tree roots; the entry itself is formed exactly as mj_jac does: u . (cdof_lin + cdof_ang x off)
        V3 u = nrm, rot = v3(0.f, 0.f, 0.f);
        if (!C6 || kd < 2) u = nrm + ld3(con + 7 + 3 * kd) * mu;
        else rot = (kd == 2 ? nrm : ld3(con + 7 + 3 * (kd - 3)));
        // (a static body has no tree root: rootidx -1, and no dof in its chain mask, so its lever arm is never used)
        const V3 off1 = cpos - ld3(p_rootcom() + 3 * max(ridx[b1], 0)), off2 = cpos - ld3(p_rootcom() + 3 * max(ridx[b2], 0));
        const int* m1 = cmask + b1 * nmw; const int* m2 = cmask + b2 * nmw;
        for (int c = 0; c < nd; c++) {
          const int d = cols.dof(c);
          const int in1 = (m1[d >> 5] >> (d & 31)) & 1, in2 = (m2[d >> 5] >> (d & 31)) & 1;
          float val = 0.f;
          if (in1 != in2) {
            const S6 cd = ld6(p_cdof() + 6 * d);
            const V3 lin = cd.l + cross(cd.a, in2 ? off2 : off1);
            val = (!C6 || kd < 2) ? dot(u, lin) : dot(nrm, lin) + mu * dot(rot, cd.a);
            if (in1) val = -val;
            vel = fmaf(val, p_qvel()[d], vel); ja = fmaf(val, p_qas()[d], ja); jw = fmaf(val, p_warm()[d], jw);
          }
          Jr[c] = val;
        }
      } else {
        const int dj = jd[info >> 4]; const float sg = (info & 1) ? -1.f : 1.f;
        for (int c = 0; c < nd; c++) Jr[c] = cols.dof(c) == dj ? sg : 0.f;
        vel = sg * p_qvel()[dj]; ja = sg * p_qas()[dj]; jw = sg * p_warm()[dj];
      }
      float aref = -B * vel - K * imp * (pos - margin);
      rowR[e] = R;
      rowb[e] = ja - aref;
      float jar = jw - aref;
      rowf[e] = jar < 0.f ? -jar / R : 0.f;
      rowres[e] = pos;    // efc_pos until the solver overwrites it with the residual (debug export reads it)
    }
    team_sync();
  }


  // ---- layout of an island's A = J M^-1 J' + R: 4x4 tiles of the lower block triangle, tile (ti, tj <= ti) at
  // 16 * (tri(ti) + tj) floats.  Row c of a tile (4 floats = one 128-bit load) sits at chunk slot c ^ ((tile >> 1) & 3), so
  // eight consecutive lanes reading the same row of eight different tiles touch eight different 16-byte bank groups.
  // Diagonal tiles are stored in full; rows / columns past n are zero (the sweep treats them as inert rows).
  __device__ __host__ __forceinline__ static int tri(int i) { return (i * (i + 1)) >> 1; }
  __device__ __host__ __forceinline__ static int a_floats(int n) { int nt = (n + 3) >> 2; return 16 * tri(nt); }
  __device__ __forceinline__ static int a_index(int i, int j) {          // requires (i >> 2) >= (j >> 2)
    int t = tri(i >> 2) + (j >> 2);
    return 16 * t + 4 * ((i & 3) ^ ((t >> 1) & 3)) + (j & 3);
  }
  // tile (this lane's rows x block b): four 128-bit shared loads; the 4x4 coefficients C[k][j] = A(4M + k, 4b + j) are
  // picked out afterwards (transposed on the fly when the tile lies above the diagonal), so the loads can be issued a
  // block ahead of their use
  __device__ __forceinline__ static void tile_raw(const float* A, int M, int b, float4 (&q)[4], bool& lower) {
    lower = M >= b; const int t = lower ? tri(M) + b : tri(b) + M; const int sw = (t >> 1) & 3;
    const float4* p = reinterpret_cast<const float4*>(A + 16 * t);
    q[0] = p[sw]; q[1] = p[1 ^ sw]; q[2] = p[2 ^ sw]; q[3] = p[3 ^ sw];
  }
  __device__ __forceinline__ static void tile_sel(const float4 (&q)[4], bool lower, float (&C)[4][4]) {
    C[0][0] = q[0].x; C[1][1] = q[1].y; C[2][2] = q[2].z; C[3][3] = q[3].w;
    C[0][1] = lower ? q[0].y : q[1].x; C[1][0] = lower ? q[1].x : q[0].y;
    C[0][2] = lower ? q[0].z : q[2].x; C[2][0] = lower ? q[2].x : q[0].z;
    C[0][3] = lower ? q[0].w : q[3].x; C[3][0] = lower ? q[3].x : q[0].w;
    C[1][2] = lower ? q[1].z : q[2].y; C[2][1] = lower ? q[2].y : q[1].z;
    C[1][3] = lower ? q[1].w : q[3].y; C[3][1] = lower ? q[3].y : q[1].w;
    C[2][3] = lower ? q[2].w : q[3].z; C[3][2] = lower ? q[3].z : q[2].w;
  }
  __device__ __forceinline__ static void tile_coeffs(const float* A, int M, int b, float (&C)[4][4]) {
    float4 q[4]; bool lower; tile_raw(A, M, b, q, lower); tile_sel(q, lower, C);
  }

  // ---- A = J M^-1 J' + R per island (mj_projectConstraint), 32 columns at a time, lane = column.
  // Phase 1 of a column block: x_j = M^-1 J_j' for the 32 columns (per-lane sparse solves on a lane-contiguous scratch);
  // phase 2: A(i, j) = J_i . x_j for all rows i >= j0, four rows at a time sharing the loads of the column.
  // Islands with more than 32 rows are built by the whole team: one warp runs phase 1, all warps split the row groups of
  // phase 2 (two team barriers per column block); small islands are built by the warp they are assigned to.
  // the column scratch is indexed by dof (32 floats per dof of the model): islands own disjoint dofs, so warps building
  // different islands never collide and the sparse solves need no column translation
  template <bool WD> __device__ __forceinline__ void build_A_solve(int k, int j0, float* scratch) {
    const int* madr = I(DI_dof_Madr); const int* ddepth = I(DI_dof_depth); const int* mcol = I(DI_Mcol);
    const int* dadr = I(DI_dof_descadr); const int* dnum = I(DI_dof_descnum); const int* dpack = I(DI_desc_pack);
    const float* LDp = p_LD(); const Cols cols = island_cols(k);
    int n = p_isl_n()[k], nd = p_isl_nd()[k], ldj = p_isl_ldj()[k];
    const float* J = xs_J<WD>(k) + jr<WD>(j0 + lane < n ? j0 + lane : 0, jblk(ldj), ldj);
    int j = j0 + lane; bool valid = j < n;
    float* x = scratch + lane;   // x[32 * dof]: lane-contiguous, conflict-free
    for (int c = 0; c < nd; c++) x[32 * cols.dof(c)] = valid ? J[c] : 0.f;
    // x <- L^-T x (gather from descendants, highest dof first), then x <- L^-1 D^-1 x (gather from ancestors):
    // per-lane sequential sparse solves with uniform control flow; the loads inside a gather are independent.
    // Columns are the island's dofs (trees in order, dofs ascending), so descending columns visit descendants first.
    for (int jj = nd - 1; jj >= 0; jj--) {
      int dj = cols.dof(jj); int dn = dnum[dj]; const int* dp = dpack + dadr[dj];
      float s0 = x[32 * dj], s1 = 0.f; int q = 0;
      for (; q + 2 <= dn; q += 2) {
        int p0 = dp[q], p1 = dp[q + 1];
        s0 = fmaf(-LDp[p0 >> 16], x[32 * (p0 & 0xffff)], s0); s1 = fmaf(-LDp[p1 >> 16], x[32 * (p1 & 0xffff)], s1);
      }
      if (q < dn) { int p0 = dp[q]; s0 = fmaf(-LDp[p0 >> 16], x[32 * (p0 & 0xffff)], s0); }
      x[32 * dj] = s0 + s1;
    }
    for (int i = 0; i < nd; i++) {
      int di = cols.dof(i); int a = madr[di], dn = ddepth[di];
      float s0 = x[32 * di] * p_invD()[di], s1 = 0.f; int m = 1;
      for (; m + 1 <= dn; m += 2) {
        s0 = fmaf(-LDp[a + m], x[32 * mcol[a + m]], s0); s1 = fmaf(-LDp[a + m + 1], x[32 * mcol[a + m + 1]], s1);
      }
      if (m <= dn) s0 = fmaf(-LDp[a + m], x[32 * mcol[a + m]], s0);
      x[32 * di] = s0 + s1;
    }
  }
  // rows i = j0 + 4 * (g0 + gs * t), t = 0, 1, ...
  __device__ __forceinline__ void build_A_dots(int k, int j0, const float* scratch, int g0, int gs) {
    int n = p_isl_n()[k], nd = p_isl_nd()[k], ldj = p_isl_ldj()[k], e0 = p_isl_adr()[k];
    const float* J = p_arena() + p_isl_J()[k]; float* A = island_A(k); const Cols cols = island_cols(k);
    int j = j0 + lane; bool valid = j < n;
    const float* x = scratch + lane;
    for (int i = j0 + 4 * g0; i < n; i += 4 * gs) {
      const float* J0 = J + i * ldj; const float* J1 = J + min(i + 1, n - 1) * ldj;
      const float* J2 = J + min(i + 2, n - 1) * ldj; const float* J3 = J + min(i + 3, n - 1) * ldj;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      if (!DYN || cols.contig) {
        const float* xd = x + 32 * cols.d0;
#pragma unroll 2
        for (int c = 0; c < nd; c++) {
          float xc = xd[32 * c];
          s0 = fmaf(J0[c], xc, s0); s1 = fmaf(J1[c], xc, s1); s2 = fmaf(J2[c], xc, s2); s3 = fmaf(J3[c], xc, s3);
        }
      } else {
#pragma unroll 2
        for (int c = 0; c < nd; c++) {
          float xc = x[32 * cols.map[c]];
          s0 = fmaf(J0[c], xc, s0); s1 = fmaf(J1[c], xc, s1); s2 = fmaf(J2[c], xc, s2); s3 = fmaf(J3[c], xc, s3);
        }
      }
      if (valid) {
        float sv[4] = {s0, s1, s2, s3};
#pragma unroll
        for (int u = 0; u < 4; u++) {
          int ii = i + u;
          if (ii < n && ii >= j) {
            float v = sv[u] + (ii == j ? p_row_R()[e0 + ii] : 0.f);
            A[a_index(ii, j)] = v;
            if ((ii >> 2) == (j >> 2) && ii != j) A[a_index(j, ii)] = v;      // diagonal tiles are stored in full
          }
        }
      }
    }
  }
  __device__ void build_A() {
    int nisl = p_misc()[MISC_NISL];
    // zero the tiles (rows / columns past n stay zero)
    for (int k = 0; k < nisl; k++) {
      int n = p_isl_n()[k]; if (!n) continue;
      bool coop = W > 1 && n > COOPMIN;
      if (!coop && p_isl_warp()[k] != wl) continue;
      float4* Az = reinterpret_cast<float4*>(island_A(k)); int nz = a_floats(n) >> 2;
      for (int i = coop ? tl : lane; i < nz; i += coop ? TEAM : 32) Az[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    team_sync();
    if (W > 1) {
      for (int k = 0; k < nisl; k++) {
        int n = p_isl_n()[k]; if (n <= COOPMIN) continue;
        float* scratch = scratch_base();
        int turn = p_isl_warp()[k];
        for (int j0 = 0; j0 < n; j0 += 32) {
          if (wl == turn) build_A_solve<false>(k, j0, scratch);
          team_sync();
          build_A_dots(k, j0, scratch, wl, W);
          team_sync();
          turn = (turn + 1 == W) ? 0 : turn + 1;
        }
      }
    }
    for (int k = 0; k < nisl; k++) {
      int n = p_isl_n()[k]; if (!n || (W > 1 && n > COOPMIN) || p_isl_warp()[k] != wl) continue;
      float* scratch = scratch_base();
      for (int j0 = 0; j0 < n; j0 += 32) { build_A_solve<false>(k, j0, scratch); sync(); build_A_dots(k, j0, scratch, 0, 1); sync(); }
    }
    team_sync();
  }

  // ---- PGS (mj_solPGS restated; row order = MuJoCo's within each island, islands are exactly decoupled so each
  // warp of the team sweeps its own islands; the stopping rule uses the improvement summed over all islands, exchanged
  // through shared memory once per iteration).  Lane L keeps rows 4L..4L+3 of the island in registers.  Rows are swept
  // four at a time: the lane that owns block b updates its four rows in order -- a chain of one FFMA + one FMNMX per row
  // -- the four force changes are broadcast with shuffles that overlap that chain, and every lane folds them into its
  // rows with the 4x4 tile (its rows x block b).  This is the same sequence of scalar row updates as mj_solPGS (same
  // order, same clamps), at a quarter of the shuffles.  The per-row state is not the residual r_k but
  //     H_k = f_k - r_k / A_kk + sum_{j<k, same block} (A_kj / A_kk) f_j ,
  // the unclamped new force of row k before the earlier rows of its block have moved, so the owner's chain starts from
  // registers: new f_0 = max(H_0, 0), new f_k = max(H_k - sum_{j<k} c_kj new f_j, 0).  A change d of block b moves
  // H_k -= (1 / A_kk) sum_j A(k, 4b + j) d_j for every other row, and for the owner only through the strict upper
  // triangle of its diagonal tile (the lower part and the diagonal are already inside H).  Raw tiles are loaded one block
  // ahead into the other of two register buffers (sweep unrolled by two, no register rotation).
  // mj_solPGS's "restore if the cost went up by more than 1e-10" guard is dropped: for these scalar row updates the cost
  // change is <= 0 in exact arithmetic, the guard only ever fires on round-off.
  struct Rows { float f[4], H[4], ainv[4], ad[4], c[6]; };
  // res[] holds H between sweeps (row_R holds 1/A_ii during the solve)
  __device__ __forceinline__ void rows_load(Rows& w, int n, int e0, const float* A, const float* res) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      int i = 4 * lane + q; bool v = i < n;
      w.f[q] = v ? p_row_f()[e0 + i] : 0.f; w.H[q] = v ? res[e0 + i] : 0.f;
      w.ad[q] = v ? A[a_index(i, i)] : 1.f; w.ainv[q] = v ? p_row_R()[e0 + i] : 1.f;
    }
    const int i0 = 4 * min(lane, ((n + 3) >> 2) - 1);         // rows past n are zero in A: their c's vanish
    w.c[0] = A[a_index(i0 + 1, i0)] * w.ainv[1]; w.c[1] = A[a_index(i0 + 2, i0)] * w.ainv[2]; w.c[2] = A[a_index(i0 + 2, i0 + 1)] * w.ainv[2];
    w.c[3] = A[a_index(i0 + 3, i0)] * w.ainv[3]; w.c[4] = A[a_index(i0 + 3, i0 + 1)] * w.ainv[3]; w.c[5] = A[a_index(i0 + 3, i0 + 2)] * w.ainv[3];
  }
  __device__ __forceinline__ void rows_store(const Rows& w, int n, int e0, float* res) {
#pragma unroll
    for (int q = 0; q < 4; q++) { int i = 4 * lane + q; if (i < n) { p_row_f()[e0 + i] = w.f[q]; res[e0 + i] = w.H[q]; } }
  }
  // q holds the raw tile of block b on entry and of block bn on exit: the coefficients are picked out first, then the same
  // registers receive the next tile while the chain below runs, so the loop body is ONE block (about 3 KB of code: the L0
  // instruction cache holds ~6 KB, and a two-block body with two named buffers did not fit -- ncu r02_a: 45 % of the active
  // warp samples in this loop were stalled on instruction fetch)
  __device__ __forceinline__ void sweep_block(Rows& w, const float* A, int M, int b, int bn, float4 (&q)[4], bool& lo, float& improvement) {
    float C[4][4]; tile_sel(q, lo, C);
    tile_raw(A, M, bn, q, lo);                                  // next block's tile: independent of the chain below
    // block owner's row updates (every lane runs them on its own registers; only lane b's are used)
    float n0 = fmaxf(w.H[0], 0.f), e0 = n0 - w.f[0];
    float d0 = __shfl_sync(B2_FULL, e0, b);
    float p1 = fmaf(-w.c[0], n0, w.H[1]), n1 = fmaxf(p1, 0.f), e1 = n1 - w.f[1];
    float d1 = __shfl_sync(B2_FULL, e1, b);
    float p2 = fmaf(-w.c[2], n1, fmaf(-w.c[1], n0, w.H[2])), n2 = fmaxf(p2, 0.f), e2 = n2 - w.f[2];
    float d2 = __shfl_sync(B2_FULL, e2, b);
    float p3 = fmaf(-w.c[5], n2, fmaf(-w.c[4], n1, fmaf(-w.c[3], n0, w.H[3]))), n3 = fmaxf(p3, 0.f), e3 = n3 - w.f[3];
    float d3 = __shfl_sync(B2_FULL, e3, b);
    const bool own = lane == b;
    // cost change of a row update: -dl (1/2 dl A_ii + residual at the time of the update), residual = (f - p) A_ii
    float ch = e0 * w.ad[0] * fmaf(0.5f, e0, w.f[0] - w.H[0]) + e1 * w.ad[1] * fmaf(0.5f, e1, w.f[1] - p1) +
               e2 * w.ad[2] * fmaf(0.5f, e2, w.f[2] - p2) + e3 * w.ad[3] * fmaf(0.5f, e3, w.f[3] - p3);
    improvement -= own ? ch : 0.f;
    w.f[0] = own ? n0 : w.f[0]; w.f[1] = own ? n1 : w.f[1]; w.f[2] = own ? n2 : w.f[2]; w.f[3] = own ? n3 : w.f[3];
    float s0 = fmaf(C[0][3], d3, fmaf(C[0][2], d2, fmaf(C[0][1], d1, own ? 0.f : C[0][0] * d0)));
    float s1 = fmaf(C[1][3], d3, fmaf(C[1][2], d2, own ? 0.f : fmaf(C[1][1], d1, C[1][0] * d0)));
    float s2 = fmaf(C[2][3], d3, own ? 0.f : fmaf(C[2][2], d2, fmaf(C[2][1], d1, C[2][0] * d0)));
    float s3 = own ? 0.f : fmaf(C[3][3], d3, fmaf(C[3][2], d2, fmaf(C[3][1], d1, C[3][0] * d0)));
    w.H[0] = fmaf(-w.ainv[0], s0, w.H[0]); w.H[1] = fmaf(-w.ainv[1], s1, w.H[1]);
    w.H[2] = fmaf(-w.ainv[2], s2, w.H[2]); w.H[3] = fmaf(-w.ainv[3], s3, w.H[3]);
  }
  __device__ __forceinline__ void sweep_island(Rows& w, int n, const float* A, float& improvement) {
    const int nb = (n + 3) >> 2; const int M = min(lane, nb - 1);
    float4 q[4]; bool lo;
    tile_raw(A, M, 0, q, lo);
#pragma unroll 1
    for (int b = 0; b < nb; b++) sweep_block(w, A, M, b, min(b + 1, nb - 1), q, lo, improvement);
  }
  __device__ void solve_pgs(unsigned long long* counters) {
    int nisl = p_misc()[MISC_NISL], iters = dim(DD_iterations);
    float scale = P.opt[DO_pgs_scale], tol = P.opt[DO_tolerance];
    float* res = p_row_res(); float* red = p_red();
    // residual r = A f + b, and warm-start acceptance: cost(f) = 1/2 f'A f + f'b > 0 -> cold start
    float cost = 0.f;
    for (int k = 0; k < nisl; k++) {
      int n = p_isl_n()[k]; if (!n || p_isl_warp()[k] != wl) continue;
      int e0 = p_isl_adr()[k]; const float* A = island_A(k); const float* fr = p_row_f() + e0;
      const int nb = (n + 3) >> 2; const int M = min(lane, nb - 1);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int b = 0; b < nb; b++) {
        float C[4][4]; tile_coeffs(A, M, b, C);
        float fb[4];
#pragma unroll
        for (int u = 0; u < 4; u++) fb[u] = (4 * b + u < n) ? fr[4 * b + u] : 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++) acc[q] = fmaf(C[q][3], fb[3], fmaf(C[q][2], fb[2], fmaf(C[q][1], fb[1], fmaf(C[q][0], fb[0], acc[q]))));
      }
      sync();
#pragma unroll
      for (int q = 0; q < 4; q++) {
        int i = 4 * lane + q;
        if (i < n) {
          float fi = fr[i], bi = p_row_b()[e0 + i]; cost += fi * (0.5f * acc[q] + bi);
          float ainv = 1.0f / A[a_index(i, i)];
          p_row_R()[e0 + i] = ainv;                             // efc_R is folded into A by now: reuse its slot for 1/A_ii
          // H_k = f_k - r_k / A_kk + sum_{j<k in block} (A_kj / A_kk) f_j
          float hk = fmaf(-(acc[q] + bi), ainv, fi);
          for (int u = 0; u < q; u++) hk = fmaf(A[a_index(i, i - q + u)] * ainv, fr[i - q + u], hk);
          res[e0 + i] = hk;
        }
      }
    }
    cost = warp_sum(cost);
    if (lane == 0) red[wl] = cost;
    team_sync();
    float total = 0.f;
#pragma unroll
    for (int q = 0; q < W; q++) total += red[q];
    if (total > 0.f) {
      for (int k = 0; k < nisl; k++) {
        int n = p_isl_n()[k], e0 = p_isl_adr()[k];
        if (p_isl_warp()[k] != wl) continue;
        for (int i = lane; i < n; i += 32) { p_row_f()[e0 + i] = 0.f; res[e0 + i] = -p_row_b()[e0 + i] * p_row_R()[e0 + i]; }   // f = 0: H = -b / A_ii
      }
    }
    sync();
    // a warp that sweeps exactly one island keeps its rows in registers across the iterations
    int mine = 0, kone = -1;
    for (int k = 0; k < nisl; k++) if (p_isl_n()[k] && p_isl_warp()[k] == wl) { mine++; kone = k; }
    const bool single = HOIST && mine == 1;
    Rows w1; int n1 = 0, e1 = 0; const float* A1 = nullptr;
    if (single) { n1 = p_isl_n()[kone]; e1 = p_isl_adr()[kone]; A1 = island_A(kone); rows_load(w1, n1, e1, A1, res); }
    int it = 0;
    for (; it < iters; it++) {
      float improvement = 0.f;
      if (single) sweep_island(w1, n1, A1, improvement);
      else {
        for (int k = 0; k < nisl; k++) {
          int n = p_isl_n()[k]; if (!n || p_isl_warp()[k] != wl) continue;
          int e0 = p_isl_adr()[k]; const float* A = island_A(k);
          Rows w; rows_load(w, n, e0, A, res);
          sweep_island(w, n, A, improvement);
          rows_store(w, n, e0, res);
        }
      }
      improvement = warp_sum(improvement);
      if (W > 1) {
        float* slot = red + 8 + (it & 1) * 4;      // double-buffered exchange
        if (lane == 0) slot[wl] = improvement;
        team_sync();
        improvement = 0.f;
#pragma unroll
        for (int q = 0; q < W; q++) improvement += slot[q];
      }
      if (improvement * scale < tol) { it++; break; }
    }
    if (single) rows_store(w1, n1, e1, res);
    if (tl == 0) { p_misc()[MISC_ITERS] = it; if (counters) atomicAdd(&counters[CTR_SOLVER_ITERS], (unsigned long long)it); }
    team_sync();
  }

  // ---- wide PGS: the same scalar row updates as solve_pgs (mj_solPGS's order within each island, islands decoupled), but
  // matrix-free.  With v = M^-1 J' f the residual of row i is J_i . v + R_i f_i + b_i and a force change d of row i moves v by
  // (M^-1 J_i') d, so a sweep needs J and B = M^-1 J' (n x nd each) instead of A (n x n): memory O(n nd), no row limit.
  // Rows are kept in blocks of four in the env's global workspace, [J rows | B rows | record] (8 ldw + 40 floats, ldw = r4(nd)),
  // and streamed through a per-warp ring in shared memory: one bulk async copy (cp.async.bulk, completion on an mbarrier) per
  // block, as many stages deep as the free arena allows (L2 latency is ~3 block times, so two stages would stall).  Lanes own dofs: the four dots J_k . v
  // are reduced by one interleaved butterfly, every lane then runs the 4-row chain on the same values
  //     f_k' = max(0, c_k f_k - e_k - ainv_k (J_k . v) - sum_{j<k} (ainv_k G_kj) d_j),   G = J_blk B_blk',  ainv = 1 / (G_kk + R_k),
  // (c_k = G_kk ainv_k, e_k = b_k ainv_k: the record), and folds the four force changes into its dofs of v.
  // one stage of the ring: an mbarrier (count 1) armed with the block's byte count, then ONE bulk copy global -> shared
  __device__ __forceinline__ static void ring_issue(float* dst, const float* src, int bytes, uint64_t* bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
  }
  __device__ __forceinline__ static void ring_wait(uint64_t* bar, unsigned parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar); uint32_t done = 0;
    while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(b), "r"(parity) : "memory");
    }
  }
  // `cnt` numbers the blocks this warp has streamed since its mbarriers were initialised: block c uses stage c % D, parity (c / D) & 1.
  // One block of look-ahead takes the warp reduction off the critical path: while the chain of block b runs, the dots of block
  // b + 1 are taken against the v that lacks block b's update, and corrected afterwards with the 4x4 tile
  // T_b = J_{b+1} B_b' from the record (s_{b+1} = J_{b+1} v_old + T_b d_b), which is the same number up to rounding.
  template <int NC>
  __device__ __forceinline__ void wide_dots(const float* blk, int ldw, const float (&v)[NC], float& s0, float& s1, float& s2, float& s3) const {
    s0 = s1 = s2 = s3 = 0.f;
#pragma unroll
    for (int q = 0; q < NC; q++) {
      int c = lane + 32 * q;
      if (c < ldw) { float vq = v[q]; s0 = fmaf(blk[c], vq, s0); s1 = fmaf(blk[ldw + c], vq, s1); s2 = fmaf(blk[2 * ldw + c], vq, s2); s3 = fmaf(blk[3 * ldw + c], vq, s3); }
    }
  }
  __device__ __forceinline__ static void wide_reduce(float& s0, float& s1, float& s2, float& s3) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s0 += __shfl_xor_sync(B2_FULL, s0, o); s1 += __shfl_xor_sync(B2_FULL, s1, o);
      s2 += __shfl_xor_sync(B2_FULL, s2, o); s3 += __shfl_xor_sync(B2_FULL, s3, o);
    }
  }
  template <int NC>
  __device__ __forceinline__ void wide_sweep(const float* g, int nb, int blkf, int ldw, float* ring, uint64_t* bars, int D, unsigned& cnt,
                                              float* f, float* vs, float& improvement) {
    float v[NC];
#pragma unroll
    for (int q = 0; q < NC; q++) { int c = lane + 32 * q; v[q] = c < ldw ? vs[c] : 0.f; }
    const int bytes = blkf * 4;
    // stage / parity of the current block, of the next one, and the stage the next copy goes to: all advanced incrementally
    // (cnt keeps (stage, parity) packed as stage | parity << 8 across sweeps, so no division by the runtime depth D is needed)
    unsigned st = cnt & 0xffu, par = cnt >> 8;
    auto adv = [D](unsigned& s_, unsigned& p_) { if (++s_ == (unsigned)D) { s_ = 0; p_ ^= 1u; } };
    unsigned si = st, pi = par;
    if (lane == 0) {
#pragma unroll 1
      for (int s = 0; s < D - 1 && s < nb; s++) { ring_issue(ring + si * blkf, g + (size_t)s * blkf, bytes, bars + si); adv(si, pi); }
    } else {
#pragma unroll 1
      for (int s = 0; s < D - 1 && s < nb; s++) adv(si, pi);
    }
    ring_wait(bars + st, par);
    float s0, s1, s2, s3;
    wide_dots<NC>(ring + st * blkf, ldw, v, s0, s1, s2, s3); wide_reduce(s0, s1, s2, s3);
    unsigned stn = st, parn = par;
#pragma unroll 1
    for (int b = 0; b < nb; b++) {
      if (b + 1 < nb) adv(stn, parn);
      ring_wait(bars + stn, parn);                   // the next block's J rows feed the look-ahead dots
      __syncwarp();                                  // every lane is done with the stage the next copy overwrites
      if (b + D - 1 < nb) { if (lane == 0) ring_issue(ring + si * blkf, g + (size_t)(b + D - 1) * blkf, bytes, bars + si); adv(si, pi); }
      const float* blk = ring + st * blkf; const float* Bb = blk + 4 * ldw;
      float p0, p1, p2, p3;
      wide_dots<NC>(ring + stn * blkf, ldw, v, p0, p1, p2, p3);
      const float4* rec = reinterpret_cast<const float4*>(blk + 8 * ldw);
      const float4 G0 = rec[0], G1 = rec[1], CC = rec[2], EE = rec[3], AI = rec[4], AD = rec[5];
      const float4 T0 = rec[6], T1 = rec[7], T2 = rec[8], T3 = rec[9];
      const float4 fo = *reinterpret_cast<const float4*>(f + 4 * b);
      wide_reduce(p0, p1, p2, p3);                   // independent of the chain below: the two interleave
      const float h0 = fmaf(-AI.x, s0, fmaf(CC.x, fo.x, -EE.x)); const float n0 = fmaxf(h0, 0.f), d0 = n0 - fo.x;
      const float h1 = fmaf(-G0.x, d0, fmaf(-AI.y, s1, fmaf(CC.y, fo.y, -EE.y))); const float n1 = fmaxf(h1, 0.f), d1 = n1 - fo.y;
      const float h2 = fmaf(-G0.z, d1, fmaf(-G0.y, d0, fmaf(-AI.z, s2, fmaf(CC.z, fo.z, -EE.z)))); const float n2 = fmaxf(h2, 0.f), d2 = n2 - fo.z;
      const float h3 = fmaf(-G1.y, d2, fmaf(-G1.x, d1, fmaf(-G0.w, d0, fmaf(-AI.w, s3, fmaf(CC.w, fo.w, -EE.w))))); const float n3 = fmaxf(h3, 0.f), d3 = n3 - fo.w;
      // cost change of a row update: d A_ii (1/2 d + f - unclamped)
      improvement -= d0 * AD.x * fmaf(0.5f, d0, fo.x - h0) + d1 * AD.y * fmaf(0.5f, d1, fo.y - h1) +
                     d2 * AD.z * fmaf(0.5f, d2, fo.z - h2) + d3 * AD.w * fmaf(0.5f, d3, fo.w - h3);
      __syncwarp();                                  // every lane has read f[4b..] before lane 0 overwrites it
      if (lane == 0) *reinterpret_cast<float4*>(f + 4 * b) = make_float4(n0, n1, n2, n3);
#pragma unroll
      for (int q = 0; q < NC; q++) {
        int c = lane + 32 * q;
        if (c < ldw) v[q] += fmaf(Bb[c], d0, Bb[ldw + c] * d1) + fmaf(Bb[2 * ldw + c], d2, Bb[3 * ldw + c] * d3);
      }
      s0 = fmaf(T0.w, d3, fmaf(T0.z, d2, fmaf(T0.y, d1, fmaf(T0.x, d0, p0))));
      s1 = fmaf(T1.w, d3, fmaf(T1.z, d2, fmaf(T1.y, d1, fmaf(T1.x, d0, p1))));
      s2 = fmaf(T2.w, d3, fmaf(T2.z, d2, fmaf(T2.y, d1, fmaf(T2.x, d0, p2))));
      s3 = fmaf(T3.w, d3, fmaf(T3.z, d2, fmaf(T3.y, d1, fmaf(T3.x, d0, p3))));
      st = stn; par = parn;
    }
    // the next sweep of this warp starts at the stage after the last block's
    adv(st, par); cnt = st | (par << 8);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < NC; q++) { int c = lane + 32 * q; if (c < ldw) vs[c] = v[q]; }
    __syncwarp();
  }
#ifdef B2_PHASE_TIMING
#define B2_WTICK(k) do { long long t_ = clock64(); if (tl == 0 && B.phase_cycles) atomicAdd(&B.phase_cycles[k], (unsigned long long)(t_ - wt0)); wt0 = t_; } while (0)
#else
#define B2_WTICK(k) do { } while (0)
#endif
  __device__ void solve_pgs_wide(unsigned long long* counters) {
#ifdef B2_PHASE_TIMING
    long long wt0 = clock64();
#endif
    const int nisl = p_misc()[MISC_NISL], iters = dim(DD_iterations), nv = dim(DD_nv), nscr = p_misc()[MISC_WSCR];
    const float scale = P.opt[DO_pgs_scale], tol = P.opt[DO_tolerance];
    float* red = p_red();
    const float* rowR = xs_row_R<true>(); const float* rowb = xs_row_b<true>(); float* rowf = xs_row_f<true>();
    // B rows and block records, 32 rows at a time; warps with a scratch set take the (island, row block) items in turn
    if (wl < nscr) {
      float* scratch = p_arena() + wl * 32 * nv; int item = 0;
      for (int k = 0; k < nisl; k++) {
        const int n = p_isl_n()[k]; if (!n) continue;
        const int nd = p_isl_nd()[k], ldj = p_isl_ldj()[k], blk = jblk(ldj), e0 = p_isl_adr()[k];
        float* Jg = xs_J<true>(k); const Cols cols = island_cols(k);
        for (int j0 = 0; j0 < n; j0 += 32, item++) {
          if (item % nscr != wl) continue;
          build_A_solve<true>(k, j0, scratch);
          sync();
          const int j = j0 + lane;
          if (j < n) {
            const float* x = scratch + lane; float* blkp = Jg + (j >> 2) * blk; float* Bj = blkp + 4 * ldj + (j & 3) * ldj;
            float G[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c = 0; c < nd; c++) {
              const float xc = x[32 * cols.dof(c)]; Bj[c] = xc;
              G[0] = fmaf(blkp[c], xc, G[0]); G[1] = fmaf(blkp[ldj + c], xc, G[1]); G[2] = fmaf(blkp[2 * ldj + c], xc, G[2]); G[3] = fmaf(blkp[3 * ldj + c], xc, G[3]);
            }
            const int q = j & 3; float* rec = blkp + 8 * ldj;
            const float gd = q == 0 ? G[0] : q == 1 ? G[1] : q == 2 ? G[2] : G[3];
            const float ad = gd + rowR[e0 + j], ainv = 1.0f / ad;
            rec[8 + q] = gd * ainv; rec[12 + q] = rowb[e0 + j] * ainv; rec[16 + q] = ainv; rec[20 + q] = ad;
            if (q == 1) rec[0] = ainv * G[0];
            if (q == 2) { rec[1] = ainv * G[0]; rec[2] = ainv * G[1]; }
            if (q == 3) { rec[3] = ainv * G[0]; rec[4] = ainv * G[1]; rec[5] = ainv * G[2]; }
            if ((j >> 2) + 1 < ((n + 3) >> 2)) {          // T[u][q] = J_{4(b+1)+u} . B_j, rows past n are zero
              const float* nxt = blkp + blk; float T[4] = {0.f, 0.f, 0.f, 0.f};
              for (int c = 0; c < nd; c++) {
                const float xc = x[32 * cols.dof(c)];
                T[0] = fmaf(nxt[c], xc, T[0]); T[1] = fmaf(nxt[ldj + c], xc, T[1]); T[2] = fmaf(nxt[2 * ldj + c], xc, T[2]); T[3] = fmaf(nxt[3 * ldj + c], xc, T[3]);
              }
              rec[24 + q] = T[0]; rec[28 + q] = T[1]; rec[32 + q] = T[2]; rec[36 + q] = T[3];
            }
          }
          sync();
        }
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");      // the blocks were written through the generic proxy; the ring reads them through the async one
    team_sync(); B2_WTICK(4);
    // f <- warm-start forces, v = B' f, cost(f) = sum_i f_i (1/2 (J_i . v + R_i f_i) + b_i); cost > 0 -> cold start
    float cost = 0.f;
    for (int k = 0; k < nisl; k++) {
      const int n = p_isl_n()[k]; if (!n || p_isl_warp()[k] != wl) continue;
      const int ldw = p_isl_ldj()[k], blk = jblk(ldw), e0 = p_isl_adr()[k], nb = (n + 3) >> 2;
      const float* Jg = xs_J<true>(k); float* f = p_arena() + p_isl_A()[k]; float* vs = f + 4 * nb;
      for (int i = lane; i < 4 * nb; i += 32) f[i] = i < n ? rowf[e0 + i] : 0.f;
      sync();
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int i = 0; i < n; i++) {
        const float fi = f[i]; if (fi == 0.f) continue;
        const float* Bi = Jg + jrow(i, blk, ldw) + 4 * ldw;
#pragma unroll
        for (int q = 0; q < 4; q++) { int c = lane + 32 * q; if (c < ldw) acc[q] = fmaf(Bi[c], fi, acc[q]); }
      }
#pragma unroll
      for (int q = 0; q < 4; q++) { int c = lane + 32 * q; if (c < ldw) vs[c] = acc[q]; }
      sync();
      for (int i = lane; i < n; i += 32) {
        const float fi = f[i]; if (fi == 0.f) continue;
        const float* Ji = Jg + jrow(i, blk, ldw); float dsum = 0.f;
        for (int c = 0; c < ldw; c++) dsum = fmaf(Ji[c], vs[c], dsum);
        cost += fi * (0.5f * fmaf(rowR[e0 + i], fi, dsum) + rowb[e0 + i]);
      }
    }
    cost = warp_sum(cost);
    if (lane == 0) red[wl] = cost;
    team_sync();
    float total = 0.f;
#pragma unroll
    for (int q = 0; q < W; q++) total += red[q];
    if (total > 0.f) {
      for (int k = 0; k < nisl; k++) {
        const int n = p_isl_n()[k]; if (!n || p_isl_warp()[k] != wl) continue;
        const int nb = (n + 3) >> 2; float* f = p_arena() + p_isl_A()[k];
        for (int i = lane; i < 4 * nb + p_isl_ldj()[k]; i += 32) f[i] = 0.f;       // f and v
      }
    }
    sync();
    // this warp's ring: D stages of one block each and their mbarriers, in the arena behind the islands' f / v
    const int D = p_misc()[MISC_WRING]; int rank = 0, sweeps = 0;
    for (int k = 0; k < nisl; k++) if (p_isl_n()[k]) { const int q = p_isl_warp()[k]; if (q == wl) sweeps = 1; }
#pragma unroll
    for (int q = 0; q < W; q++) { bool any = false; for (int k = 0; k < nisl; k++) any |= p_isl_n()[k] && p_isl_warp()[k] == q; if (q < wl && any) rank++; }
    float* ring = p_arena() + p_misc()[MISC_WRB] + rank * D * (wide_blkf_max() + 4);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + D * wide_blkf_max());
    unsigned cnt = 0;
    if (sweeps && lane == 0) {
      for (int st = 0; st < D; st++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bars + st)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp(); B2_WTICK(5);
#ifdef B2_PHASE_TIMING
    if (tl == 0 && B.phase_cycles) { atomicAdd(&B.phase_cycles[14], (unsigned long long)D); atomicAdd(&B.phase_cycles[15], 1ull); atomicAdd(&B.phase_cycles[8], (unsigned long long)p_misc()[MISC_NEFC]); }
#endif
    int it = 0;
    for (; it < iters; it++) {
      float improvement = 0.f;
      for (int k = 0; k < nisl; k++) {
        const int n = p_isl_n()[k]; if (!n || p_isl_warp()[k] != wl) continue;
        const int ldw = p_isl_ldj()[k], blkf = jblk(ldw), nb = (n + 3) >> 2;
        const float* Jg = xs_J<true>(k); float* f = p_arena() + p_isl_A()[k]; float* vs = f + 4 * nb;
        if (ldw <= 32) wide_sweep<1>(Jg, nb, blkf, ldw, ring, bars, D, cnt, f, vs, improvement);
        else if (ldw <= 64) wide_sweep<2>(Jg, nb, blkf, ldw, ring, bars, D, cnt, f, vs, improvement);
        else wide_sweep<4>(Jg, nb, blkf, ldw, ring, bars, D, cnt, f, vs, improvement);
      }
      if (W > 1) {
        float* slot = red + 8 + (it & 1) * 4;      // double-buffered exchange (the sweeps leave `improvement` warp-uniform)
        if (lane == 0) slot[wl] = improvement;
#ifdef B2_PHASE_TIMING
        long long tb0 = clock64();
#endif
        team_sync();
#ifdef B2_PHASE_TIMING
        if (tl == 0 && B.phase_cycles) atomicAdd(&B.phase_cycles[18], (unsigned long long)(clock64() - tb0));
#endif
        improvement = 0.f;
#pragma unroll
        for (int q = 0; q < W; q++) improvement += slot[q];
      }
      if (improvement * scale < tol) { it++; break; }
    }
    B2_WTICK(6);
#ifdef B2_PHASE_TIMING
    if (tl == 0 && B.phase_cycles) atomicAdd(&B.phase_cycles[7], (unsigned long long)it);
#endif
    if (sweeps && lane == 0) {      // the arena is re-carved every pass: retire the mbarriers before their words are reused
      for (int st = 0; st < D; st++) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bars + st)) : "memory");
    }
    __syncwarp();
    for (int k = 0; k < nisl; k++) {
      const int n = p_isl_n()[k]; if (!n || p_isl_warp()[k] != wl) continue;
      const int e0 = p_isl_adr()[k]; const float* f = p_arena() + p_isl_A()[k];
      for (int i = lane; i < n; i += 32) rowf[e0 + i] = f[i];
    }
    if (tl == 0) { p_misc()[MISC_ITERS] = it; if (counters) atomicAdd(&counters[CTR_SOLVER_ITERS], (unsigned long long)it); }
    __threadfence_block();
    team_sync();
  }
  // pass 1 of a forward evaluation in the wide tier (cold: a few per cent of the passes; kept out of line)
  __device__ __forceinline__ void wide_pass(unsigned long long* counters) {
    if (p_misc()[MISC_NEFC] > 0) {
      fill_rows<true>();
      asm volatile("fence.proxy.async;" ::: "memory");
      team_sync();
      if (newton()) solve_newton<true>(counters);
      else solve_pgs_wide(counters);
    }
    else if (tl == 0) p_misc()[MISC_ITERS] = 0;
    if (wl == 0) qfrc_constraint<true>(newton());
  }

  // ---- Newton solver (mj_solNewton restated for the one-sided quadratic rows of limits and pyramidal contacts), one warp
  // per island:  minimise  1/2 (a - a_s)' M (a - a_s) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2  over the island's
  // accelerations with the exact Hessian H = M + J' diag(D active) J (dense Cholesky in shared memory, nd <= 64) and an
  // exact line search on the piecewise-quadratic cost (safeguarded Newton on its derivative).  The optimum is unique, so
  // parity with the fp64 oracle is on the converged solution, not on the iteration path.  Leaves efc_force in row_f.
#ifndef B2_NEWTON_SPARSE128
#define B2_NEWTON_SPARSE128 0   // sparse Hessian build for islands of 65-128 dofs: measured no gain on construction (r02), kept as a switch
#endif
#ifndef B2_NEWTON_RTOL
#define B2_NEWTON_RTOL 1e-5f
#endif
#ifndef B2_NEWTON_AFLOOR
#define B2_NEWTON_AFLOOR 10.0f
#endif
#ifndef B2_NEWTON_NOISE
#define B2_NEWTON_NOISE 5e-6f
#endif
#ifndef B2_NEWTON_GRADNOISE
#define B2_NEWTON_GRADNOISE 1e-5f
#endif
  // H and a dense copy of the island's block of M (packed lower triangles: multiplying by the sparse ancestor rows of M needs
  // shared-memory atomics), nine vectors of nd, jv
  __host__ __device__ __forceinline__ static int newton_floats(int n, int nd) { return 2 * r4(nd * (nd + 1) / 2) + 9 * r4(nd) + r4(n) + 8; }
  // ---- team-wide Newton solve of one island (same mathematics and memory layout as the one-warp loop in solve_newton below).
  // An env of the Newton tasks has one big island (the humanoid or the arm with whatever it touches, 30-99 dofs) beside a few
  // 6-dof free bodies; with one warp per island two of the team's three warps sat at the closing barrier for 80 % of the pass
  // (r02 profiles: stalled_barrier 7.7-11.7 per issue).  Here every phase of the iteration is spread over the team's threads:
  // rows of J over threads, the Hessian's column pairs of a contact group over threads (each warp compacts the group's column
  // list for itself), the Cholesky trailing update as rows over lanes x columns over warps, M s on warp 0 while the other
  // warps form J s.  All exit tests use team-reduced values summed in a fixed order, so the control flow is team-uniform and
  // the result does not depend on timing.  The triangular solves stay on warp 0 (a dependent chain).
  // sums of two values over the team, same result in every thread; alternates between the two halves of `red` so that one
  // barrier per call is enough (a half is rewritten only after the barrier of the following call)
  __device__ __forceinline__ void team_sum2(float& x, float& y2, int& par) {
    warp_sum2(x, y2);
    float* r = p_red() + 8 * par; par ^= 1;
    if (lane == 0) { r[2 * wl] = x; r[2 * wl + 1] = y2; }
    team_sync();
    x = r[0]; y2 = r[1];
#pragma unroll
    for (int q = 1; q < W; q++) { x += r[2 * q]; y2 += r[2 * q + 1]; }
  }
  __device__ __forceinline__ float team_max(float x, int& par) {
#pragma unroll
    for (int o = 16; o; o >>= 1) x = fmaxf(x, __shfl_xor_sync(B2_FULL, x, o));
    float* r = p_red() + 8 * par; par ^= 1;
    if (lane == 0) r[2 * wl] = x;
    team_sync();
    x = r[0];
#pragma unroll
    for (int q = 1; q < W; q++) x = fmaxf(x, r[2 * q]);
    return x;
  }
  // y = Md x on one warp (lanes over columns); closes with a warp barrier
  __device__ __forceinline__ void warp_mulMd(const float* Md, int nd, const float* x, float* y) {
    for (int c = lane; c < nd; c += 32) {
      const float* Mc = Md + c * (c + 1) / 2; float s0 = 0.f, s1 = 0.f;
      for (int e = 0; e <= c; e++) s0 = fmaf(Mc[e], x[e], s0);
      for (int e = c + 1; e < nd; e++) s1 = fmaf(Md[e * (e + 1) / 2 + c], x[e], s1);
      y[c] = s0 + s1;
    }
    sync();
  }
  // y = Md x for a dense packed lower triangle (one column per thread; the caller places the barriers)
  __device__ __forceinline__ void team_mulMd(const float* Md, int nd, const float* x, float* y) {
    for (int c = tl; c < nd; c += TEAM) {
      const float* Mc = Md + c * (c + 1) / 2; float s0 = 0.f, s1 = 0.f;
      int e = 0;
      for (; e + 1 <= c; e += 2) { s0 = fmaf(Mc[e], x[e], s0); s1 = fmaf(Mc[e + 1], x[e + 1], s1); }
      if (e <= c) s0 = fmaf(Mc[e], x[e], s0);
      for (e = c + 1; e < nd; e++) s1 = fmaf(Md[e * (e + 1) / 2 + c], x[e], s1);
      y[c] = s0 + s1;
    }
  }
  template <bool WD> __device__ __forceinline__ int newton_island_team(int k, unsigned long long* counters) {
    const int* madr = I(DI_dof_Madr); const int* ddepth = I(DI_dof_depth); const int* mcol = I(DI_Mcol);
    const int iters = dim(DD_iterations);
    const float scale = P.opt[DO_pgs_scale], tol = P.opt[DO_tolerance];
    const int n = p_isl_n()[k];
    const int nd = p_isl_nd()[k], ndp = r4(nd), ldj = p_isl_ldj()[k], e0 = p_isl_adr()[k], nh = nd * (nd + 1) / 2;
#ifdef B2_PHASE_TIMING
    const long long tn0 = clock64(); long long tq = tn0;
#define B2_NTT(slot) do { if (tl == 0 && B.phase_cycles) { const long long t_ = clock64(); atomicAdd(&B.phase_cycles[slot], (unsigned long long)(t_ - tq)); tq = t_; } } while (0)
#else
#define B2_NTT(slot) do { } while (0)
#endif
    const float* J = xs_J<WD>(k); const Cols cols = island_cols(k);
    float* H = island_A(k); float* a = H + r4(nh); float* as = a + ndp; float* fs = as + ndp;
    float* Ma = fs + ndp; float* grad = Ma + ndp; float* srch = grad + ndp; float* Mv = srch + ndp; float* y = Mv + ndp;
    float* wrm = y + ndp; float* jv = WD ? xs_J<WD>(k) + r4(n * ldj) : wrm + ndp;
    float* Md = wrm + ndp + (WD ? 0 : r4(n));      // dense packed copy of the island's block of M
    float* Dr = xs_row_R<WD>() + e0; float* aref = xs_row_res<WD>() + e0; float* jar = xs_row_f<WD>() + e0; const float* bb = xs_row_b<WD>() + e0;
    int par = 0;
    for (int c = tl; c < nd; c += TEAM) { int d = cols.dof(c); as[c] = p_qas()[d]; fs[c] = p_qfs()[d]; wrm[c] = p_warm()[d]; }
    for (int q = tl; q < nh; q += TEAM) Md[q] = 0.f;
    team_sync();
    for (int c = tl; c < nd; c += TEAM) {
      int d = cols.dof(c), m0 = madr[d], dep = ddepth[d];
      for (int u = 0; u <= dep; u++) { int ca = p_dof_col()[mcol[m0 + u]]; int r = max(c, ca), cc = min(c, ca); Md[r * (r + 1) / 2 + cc] = p_M()[m0 + u]; }
    }
    for (int i = tl; i < n; i += TEAM) {
      const float* Ji = J + i * ldj; float s = 0.f;
      for (int c = 0; c < nd; c++) s = fmaf(Ji[c], as[c], s);
      aref[i] = s - bb[i]; Dr[i] = fminf(1.0f / Dr[i], 1e8f);
    }
    for (int c = tl; c < nd; c += TEAM) y[c] = wrm[c] - as[c];
    team_sync();
    team_mulMd(Md, nd, y, Mv);
    team_sync();
    float cw = 0.f, cs = 0.f;
    for (int i = tl; i < n; i += TEAM) {
      const float* Ji = J + i * ldj; float sw = 0.f, ss = 0.f;
      for (int c = 0; c < nd; c++) { sw = fmaf(Ji[c], wrm[c], sw); ss = fmaf(Ji[c], as[c], ss); }
      sw -= aref[i]; ss -= aref[i];
      if (sw < 0.f) cw += 0.5f * Dr[i] * sw * sw;
      if (ss < 0.f) cs += 0.5f * Dr[i] * ss * ss;
    }
    for (int c = tl; c < nd; c += TEAM) cw += 0.5f * Mv[c] * y[c];
    team_sum2(cw, cs, par);
    for (int c = tl; c < nd; c += TEAM) a[c] = (cw < cs) ? wrm[c] : as[c];
    team_sync();
    team_mulMd(Md, nd, a, Ma);
    team_sync();
    B2_NTT(22);
    int it = 0;
    for (; it < iters; it++) {
      for (int i = tl; i < n; i += TEAM) {
        const float* Ji = J + i * ldj; float s = 0.f;
        for (int c = 0; c < nd; c++) s = fmaf(Ji[c], a[c], s);
        s -= aref[i];
        jar[i] = s; jv[i] = s < 0.f ? Dr[i] : 0.f;      // D of the active rows, zero for the others (jv is free until the search direction exists)
      }
      team_sync();
      float g2 = 0.f, nres = 0.f;
      for (int c = tl; c < nd; c += TEAM) {
        float g = Ma[c] - fs[c], sabs = fabsf(Ma[c]) + fabsf(fs[c]);
        const float* Jc = J + c;
#pragma unroll 4
        for (int i = 0; i < n; i++) { const float t = Jc[i * ldj] * (jv[i] * jar[i]); g += t; sabs += fabsf(t); }
        grad[c] = g; g2 = fmaf(g, g, g2);
        if (fabsf(g) > B2_NEWTON_GRADNOISE * sabs) nres += 1.f;
      }
      team_sum2(g2, nres, par);
      B2_NTT(23);
      if (scale * sqrtf(g2) < tol) break;
      if (nres == 0.f) break;      // every gradient entry below the rounding noise of its own terms (see the one-warp loop)
      // H = M + J' diag(D active) J, lower triangle, one 4 x 4 tile per thread: a row of J contributes sixteen multiply-adds for
      // nine shared-memory loads (four entries of each of the tile's two column groups and the row's D), where one entry per
      // thread paid three loads per multiply-add and was bound by shared-memory bandwidth.  jv holds D of the active rows and
      // zero for the others, so the inner loop has no branch.
      {
        const int nt = (nd + 3) >> 2, ntile = nt * (nt + 1) / 2;
        for (int q = tl; q < ntile; q += TEAM) {
          int tr = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
          while (tr * (tr + 1) / 2 > q) tr--;
          while ((tr + 1) * (tr + 2) / 2 <= q) tr++;
          const int tc = q - tr * (tr + 1) / 2, r0 = 4 * tr, c0 = 4 * tc;
          // columns past nd read the row's last valid column and are never stored
          const int ra = min(r0, nd - 1), rb = min(r0 + 1, nd - 1), rc = min(r0 + 2, nd - 1), rd = min(r0 + 3, nd - 1);
          const int ca = min(c0, nd - 1), cb = min(c0 + 1, nd - 1), cc = min(c0 + 2, nd - 1), cd = min(c0 + 3, nd - 1);
          float h[4][4];
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int v = 0; v < 4; v++) h[u][v] = 0.f;
#pragma unroll 2
          for (int i = 0; i < n; i++) {
            const float* Ji = J + i * ldj; const float w = jv[i];
            const float a0 = Ji[ra] * w, a1 = Ji[rb] * w, a2 = Ji[rc] * w, a3 = Ji[rd] * w;
            const float b0 = Ji[ca], b1 = Ji[cb], b2 = Ji[cc], b3 = Ji[cd];
            h[0][0] = fmaf(a0, b0, h[0][0]); h[0][1] = fmaf(a0, b1, h[0][1]); h[0][2] = fmaf(a0, b2, h[0][2]); h[0][3] = fmaf(a0, b3, h[0][3]);
            h[1][0] = fmaf(a1, b0, h[1][0]); h[1][1] = fmaf(a1, b1, h[1][1]); h[1][2] = fmaf(a1, b2, h[1][2]); h[1][3] = fmaf(a1, b3, h[1][3]);
            h[2][0] = fmaf(a2, b0, h[2][0]); h[2][1] = fmaf(a2, b1, h[2][1]); h[2][2] = fmaf(a2, b2, h[2][2]); h[2][3] = fmaf(a2, b3, h[2][3]);
            h[3][0] = fmaf(a3, b0, h[3][0]); h[3][1] = fmaf(a3, b1, h[3][1]); h[3][2] = fmaf(a3, b2, h[3][2]); h[3][3] = fmaf(a3, b3, h[3][3]);
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int r = r0 + u;
            if (r < nd) {
#pragma unroll
              for (int v = 0; v < 4; v++) { const int c = c0 + v; if (c <= r) { const int x = r * (r + 1) / 2 + c; H[x] = h[u][v] + Md[x]; } }
            }
          }
        }
        team_sync();
      }
      B2_NTT(24);
      for (int c = tl; c < nd; c += TEAM) Mv[c] = rsqrtf(fmaxf(H[c * (c + 1) / 2 + c], 1e-30f));   // Mv doubles as the scale vector here
      team_sync();
      for (int r = wl; r < nd; r += W) { float sr = Mv[r]; float* Hr = H + r * (r + 1) / 2; for (int c = lane; c <= r; c += 32) Hr[c] *= sr * Mv[c]; }
      team_sync();
      B2_NTT(25);
      // Cholesky H = L L' in place, right-looking in panels of four pivots: every thread factors the panel's 4 x 4 diagonal block
      // for itself (ten broadcast loads), rows below the panel are solved against it one row per thread in registers, and the
      // trailing update subtracts four rank-one terms per entry (rows over lanes, columns over warps).  Two barriers per panel
      // instead of two per pivot; reciprocal pivots go to wrm as in the one-warp loop.
      for (int p0 = 0; p0 < nd; p0 += 4) {
        const int b = min(4, nd - p0);
        const float* D0 = H + p0 * (p0 + 1) / 2 + p0; const float* D1 = D0 + p0 + 1; const float* D2 = D1 + p0 + 2; const float* D3 = D2 + p0 + 3;
        const float i00 = rsqrtf(fmaxf(D0[0], 1e-7f));
        float l10 = 0.f, l20 = 0.f, l21 = 0.f, l30 = 0.f, l31 = 0.f, l32 = 0.f, i11 = 1.f, i22 = 1.f, i33 = 1.f;
        if (b > 1) { l10 = D1[0] * i00; i11 = rsqrtf(fmaxf(fmaf(-l10, l10, D1[1]), 1e-7f)); }
        if (b > 2) { l20 = D2[0] * i00; l21 = fmaf(-l20, l10, D2[1]) * i11; i22 = rsqrtf(fmaxf(fmaf(-l21, l21, fmaf(-l20, l20, D2[2])), 1e-7f)); }
        if (b > 3) {
          l30 = D3[0] * i00; l31 = fmaf(-l30, l10, D3[1]) * i11; l32 = fmaf(-l31, l21, fmaf(-l30, l20, D3[2])) * i22;
          i33 = rsqrtf(fmaxf(fmaf(-l32, l32, fmaf(-l31, l31, fmaf(-l30, l30, D3[3]))), 1e-7f));
        }
        for (int i = p0 + b + tl; i < nd; i += TEAM) {
          float* Hi = H + i * (i + 1) / 2 + p0;
          const float x0 = Hi[0] * i00; Hi[0] = x0;
          if (b > 1) { const float x1 = fmaf(-x0, l10, Hi[1]) * i11; Hi[1] = x1;
            if (b > 2) { const float x2 = fmaf(-x1, l21, fmaf(-x0, l20, Hi[2])) * i22; Hi[2] = x2;
              if (b > 3) { Hi[3] = fmaf(-x2, l32, fmaf(-x1, l31, fmaf(-x0, l30, Hi[3]))) * i33; } } }
        }
        team_sync();      // panel rows are final, and everyone has read the diagonal block: its factor may replace it now
        if (tl == 0) {
          wrm[p0] = i00;
          if (b > 1) { wrm[p0 + 1] = i11; const_cast<float*>(D1)[0] = l10; }
          if (b > 2) { wrm[p0 + 2] = i22; const_cast<float*>(D2)[0] = l20; const_cast<float*>(D2)[1] = l21; }
          if (b > 3) { wrm[p0 + 3] = i33; const_cast<float*>(D3)[0] = l30; const_cast<float*>(D3)[1] = l31; const_cast<float*>(D3)[2] = l32; }
        }
        if (b == 4) {
          for (int i = p0 + 4 + lane; i < nd; i += 32) {
            float* Hi = H + i * (i + 1) / 2; const float a0 = Hi[p0], a1 = Hi[p0 + 1], a2 = Hi[p0 + 2], a3 = Hi[p0 + 3];
#pragma unroll 2
            for (int c = p0 + 4 + wl; c <= i; c += W) {
              const float* Hc = H + c * (c + 1) / 2 + p0;
              Hi[c] = fmaf(-a3, Hc[3], fmaf(-a2, Hc[2], fmaf(-a1, Hc[1], fmaf(-a0, Hc[0], Hi[c]))));
            }
          }
        }
        team_sync();      // the trailing update reaches the next panel's diagonal block (a shorter last panel has no rows below it)
      }
      B2_NTT(26);
      if (wl == 0) {
        // search = -H^-1 grad on warp 0, unknowns in registers (lane owns rows lane, lane + 32, ...): a step is one shuffle,
        // one multiply by the reciprocal pivot and the lane's own multiply-adds; the column of L is loaded ahead of the chain
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = lane + 32 * u; v[u] = i < nd ? -grad[i] * Mv[i] : 0.f; }
#pragma unroll
        for (int ub = 0; ub < 4; ub++) {
          if (32 * ub < nd) {
            const int rend = min(32, nd - 32 * ub);
            for (int rr = 0; rr < rend; rr++) {
              const int r = 32 * ub + rr;
              float hcol[4];
#pragma unroll
              for (int u = 0; u < 4; u++) { const int i = lane + 32 * u; hcol[u] = (u >= ub && i > r && i < nd) ? H[i * (i + 1) / 2 + r] : 0.f; }
              const float zr = __shfl_sync(B2_FULL, v[ub], rr) * wrm[r];
              if (lane == rr) v[ub] = zr;
#pragma unroll
              for (int u = 0; u < 4; u++) if (u >= ub) v[u] = fmaf(-hcol[u], zr, v[u]);
            }
          }
        }
#pragma unroll
        for (int ub = 3; ub >= 0; ub--) {
          if (32 * ub < nd) {
            const int rend = min(32, nd - 32 * ub);
            for (int rr = rend - 1; rr >= 0; rr--) {
              const int r = 32 * ub + rr; const float* Hr = H + r * (r + 1) / 2;
              float hrow[4];
#pragma unroll
              for (int u = 0; u < 4; u++) { const int i = lane + 32 * u; hrow[u] = (u <= ub && i < r) ? Hr[i] : 0.f; }
              const float xr = __shfl_sync(B2_FULL, v[ub], rr) * wrm[r];
              if (lane == rr) v[ub] = xr;
#pragma unroll
              for (int u = 0; u < 4; u++) if (u <= ub) v[u] = fmaf(-hrow[u], xr, v[u]);
            }
          }
        }
        // undo the scaling, s = S (S H S)^-1 S (-g); y <- fp32 noise floor of each acceleration, (|f_smooth| + |M a|) / H_cc
#pragma unroll
        for (int u = 0; u < 4; u++) { const int c = lane + 32 * u; if (c < nd) { const float sc = Mv[c]; srch[c] = v[u] * sc; y[c] = sc * sc * (fabsf(fs[c]) + fabsf(Ma[c])); } }
      }
      team_sync();
      B2_NTT(27);
      // M s and J s, then the two quadratic coefficients of the smooth part
      team_mulMd(Md, nd, srch, Mv);
      for (int i = tl; i < n; i += TEAM) {
        const float* Ji = J + i * ldj; float sj = 0.f;
        for (int c = 0; c < nd; c++) sj = fmaf(Ji[c], srch[c], sj);
        jv[i] = sj;
      }
      float q1 = 0.f, q2 = 0.f;
      for (int c = tl; c < nd; c += TEAM) { q1 = fmaf(srch[c], Ma[c] - fs[c], q1); q2 = fmaf(srch[c], Mv[c], q2); }      // own column of Mv
      team_sum2(q1, q2, par);
      B2_NTT(28);
      // exact minimisation along the search direction (same safeguarded Newton on the slope as the one-warp loop)
      float alpha = 0.f, lo = 0.f, hi = -1.f;
      for (int ls = 0; ls < 40; ls++) {
        float d1 = 0.f, d2 = 0.f, dn = 0.f;
        for (int i = tl; i < n; i += TEAM) { float x = fmaf(alpha, jv[i], jar[i]); if (x < 0.f) { float t = Dr[i] * jv[i]; d1 = fmaf(t, x, d1); d2 = fmaf(t, jv[i], d2); dn += fabsf(t * x); } }
        team_sum2(d1, d2, par);
        if (ls == 0) { float z = 0.f; team_sum2(dn, z, par); }
        d1 += q1 + alpha * q2; d2 += q2;
        if (ls == 0 && fabsf(d1) <= 2.4e-7f * (dn + fabsf(q1))) { alpha = 1.0f; break; }
        if (fabsf(d1) < 1e-6f * (1.0f + fabsf(q1))) break;
        if (d1 < 0.f) lo = alpha; else hi = alpha;
        if (!(d2 > 0.f)) break;
        float na = alpha - d1 / d2;
        if (hi > 0.f && (na <= lo || na >= hi)) na = 0.5f * (lo + hi);
        if (na < 0.f) na = 0.f;
        bool done = fabsf(na - alpha) < 1e-7f * (1.0f + fabsf(alpha));
        alpha = na;
        if (done) break;
      }
      B2_NTT(29);
      float amx = B2_NEWTON_AFLOOR;
      for (int c = tl; c < nd; c += TEAM) { a[c] = fmaf(alpha, srch[c], a[c]); Ma[c] = fmaf(alpha, Mv[c], Ma[c]); amx = fmaxf(amx, fmaxf(fabsf(a[c]), fabsf(as[c]))); }
      amx = team_max(amx, par);
      float nmov = 0.f, z = 0.f;
      for (int c = tl; c < nd; c += TEAM) if (fabsf(alpha * srch[c]) > fmaf(B2_NEWTON_RTOL, amx, B2_NEWTON_NOISE * y[c])) nmov += 1.f;
      team_sum2(nmov, z, par);
      B2_NTT(30);
      if (nmov == 0.f) { it++; break; }
    }
    team_sync();
    for (int i = tl; i < n; i += TEAM) {
      const float* Ji = J + i * ldj; float s = 0.f;
      for (int c = 0; c < nd; c++) s = fmaf(Ji[c], a[c], s);
      s -= aref[i];
      jar[i] = s < 0.f ? -Dr[i] * s : 0.f;
    }
    for (int c = tl; c < nd; c += TEAM) y[c] = a[c] - as[c];
    team_sync();
    team_mulMd(Md, nd, y, Mv);
    for (int c = tl; c < nd; c += TEAM) { int d = cols.dof(c); p_qfc()[d] = Mv[c]; p_qacc()[d] = Mv[c]; }      // own column of Mv
    if (it >= iters && tl == 0 && counters) atomicAdd(&counters[CTR_ARENA_SPILL], 1ull);
    B2_NTT(31);
#ifdef B2_PHASE_TIMING
    if (tl == 0 && B.phase_cycles) { atomicAdd(&B.phase_cycles[20], 1ull); atomicAdd(&B.phase_cycles[21], (unsigned long long)it); }
#endif
    return it;
  }
  // ---- Newton solve of a free body's island (<= 6 dofs, <= 32 rows) in registers: lane i owns row i of J, the island's
  // vectors are replicated in every lane, sums over rows are warp butterflies, and the 6 x 6 Hessian is factored by every lane
  // for itself.  Same mathematics and exit tests as the loop in solve_newton; no shared-memory round trips or warp barriers
  // inside the iteration (the generic loop spends ~30 k cycles on such an island, nearly all of it latency between its many
  // short phases).  Shared memory holds only the dense block of M (in the island's H slot) and the rows' inputs / outputs.
  template <bool WD> __device__ __forceinline__ int newton_island_small(int k, unsigned long long* counters) {
    constexpr int ND = 6;
    const int* madr = I(DI_dof_Madr); const int* ddepth = I(DI_dof_depth); const int* mcol = I(DI_Mcol);
    const int iters = dim(DD_iterations);
    const float scale = P.opt[DO_pgs_scale], tol = P.opt[DO_tolerance];
    const int n = p_isl_n()[k], nd = p_isl_nd()[k], ldj = p_isl_ldj()[k], e0 = p_isl_adr()[k], nh = nd * (nd + 1) / 2;
    const float* J = xs_J<WD>(k); const Cols cols = island_cols(k);
    float* Md = island_A(k);
    float* Dr = xs_row_R<WD>() + e0; float* jar = xs_row_f<WD>() + e0; const float* bb = xs_row_b<WD>() + e0;
    const bool row = lane < n;
    for (int q = lane; q < nh; q += 32) Md[q] = 0.f;
    sync();
    if (lane < nd) {
      int d = cols.dof(lane), m0 = madr[d], dep = ddepth[d];
      for (int u = 0; u <= dep; u++) { int ca = p_dof_col()[mcol[m0 + u]]; int r = max(lane, ca), cc = min(lane, ca); Md[r * (r + 1) / 2 + cc] = p_M()[m0 + u]; }
    }
    float Jr[ND], as[ND], fs[ND], a[ND], Ma[ND], wr[ND];
#pragma unroll
    for (int c = 0; c < ND; c++) {
      const bool on = c < nd; const int d = on ? cols.dof(c) : 0;
      Jr[c] = (row && on) ? J[lane * ldj + c] : 0.f;
      as[c] = on ? p_qas()[d] : 0.f; fs[c] = on ? p_qfs()[d] : 0.f; wr[c] = on ? p_warm()[d] : 0.f;
    }
    const float Dl = row ? fminf(1.0f / Dr[lane], 1e8f) : 0.f;
    sync();
    // y = Md x: lane c forms entry c, then every lane collects all of them
    auto mulM = [&](const float (&x)[ND], float (&y)[ND]) {
      float yc = 0.f;
      if (lane < nd) {
#pragma unroll
        for (int e = 0; e < ND; e++) if (e < nd) { const int r = max(lane, e), cc = min(lane, e); yc = fmaf(Md[r * (r + 1) / 2 + cc], x[e], yc); }
      }
#pragma unroll
      for (int c = 0; c < ND; c++) y[c] = __shfl_sync(B2_FULL, yc, c);
    };
    auto rowdot = [&](const float (&x)[ND]) { float s = 0.f;
#pragma unroll
      for (int c = 0; c < ND; c++) s = fmaf(Jr[c], x[c], s);
      return s; };
    const float aref = rowdot(as) - (row ? bb[lane] : 0.f);
    {
      float y[ND], My[ND];
#pragma unroll
      for (int c = 0; c < ND; c++) y[c] = wr[c] - as[c];
      mulM(y, My);
      const float sw = rowdot(wr) - aref, ss = rowdot(as) - aref;
      float cw = (row && sw < 0.f) ? 0.5f * Dl * sw * sw : 0.f, cs = (row && ss < 0.f) ? 0.5f * Dl * ss * ss : 0.f;
      warp_sum2(cw, cs);
#pragma unroll
      for (int c = 0; c < ND; c++) cw = fmaf(0.5f * My[c], y[c], cw);
#pragma unroll
      for (int c = 0; c < ND; c++) a[c] = (cw < cs) ? wr[c] : as[c];
    }
    mulM(a, Ma);
    float jr = 0.f;
    int it = 0;
    for (; it < iters; it++) {
      jr = rowdot(a) - aref;
      const bool act = row && jr < 0.f;
      const float t = act ? Dl * jr : 0.f, w = act ? Dl : 0.f;
      float grad[ND]; float g2 = 0.f; bool resolved = false;
#pragma unroll
      for (int c = 0; c < ND; c++) {
        float g = Jr[c] * t, sa = fabsf(g);
        warp_sum2(g, sa);
        g += Ma[c] - fs[c]; sa += fabsf(Ma[c]) + fabsf(fs[c]);
        grad[c] = g; g2 = fmaf(g, g, g2);
        resolved |= c < nd && fabsf(g) > B2_NEWTON_GRADNOISE * sa;
      }
      if (scale * sqrtf(g2) < tol) break;
      if (!resolved) break;
      // H = M + J' diag(D active) J (all lanes hold all of it), scaled to a unit diagonal, Cholesky in registers
      float H[ND][ND], sc[ND], ip[ND];
#pragma unroll
      for (int r = 0; r < ND; r++) {
#pragma unroll
        for (int c = 0; c <= r; c++) {
          float h = 0.f;
          if (r < nd) { h = w * Jr[r] * Jr[c]; h = warp_sum(h) + Md[r * (r + 1) / 2 + c]; }
          H[r][c] = h;
        }
      }
#pragma unroll
      for (int c = 0; c < ND; c++) sc[c] = c < nd ? rsqrtf(fmaxf(H[c][c], 1e-30f)) : 1.f;
#pragma unroll
      for (int r = 0; r < ND; r++) {
#pragma unroll
        for (int c = 0; c <= r; c++) H[r][c] = r < nd ? H[r][c] * sc[r] * sc[c] : (r == c ? 1.f : 0.f);
      }
#pragma unroll
      for (int p = 0; p < ND; p++) {
        const float inv = rsqrtf(fmaxf(H[p][p], 1e-7f)); ip[p] = inv;
#pragma unroll
        for (int i = p + 1; i < ND; i++) H[i][p] *= inv;
#pragma unroll
        for (int i = p + 1; i < ND; i++) {
#pragma unroll
          for (int c = p + 1; c <= i; c++) H[i][c] = fmaf(-H[i][p], H[c][p], H[i][c]);
        }
      }
      float x[ND];
#pragma unroll
      for (int c = 0; c < ND; c++) x[c] = -grad[c] * sc[c];
#pragma unroll
      for (int r = 0; r < ND; r++) {
        x[r] *= ip[r];
#pragma unroll
        for (int i = r + 1; i < ND; i++) x[i] = fmaf(-H[i][r], x[r], x[i]);
      }
#pragma unroll
      for (int r = ND - 1; r >= 0; r--) {
        x[r] *= ip[r];
#pragma unroll
        for (int i = 0; i < r; i++) x[i] = fmaf(-H[r][i], x[r], x[i]);
      }
      float srch[ND], Mv[ND], noise[ND];
#pragma unroll
      for (int c = 0; c < ND; c++) { srch[c] = c < nd ? x[c] * sc[c] : 0.f; noise[c] = sc[c] * sc[c] * (fabsf(fs[c]) + fabsf(Ma[c])); }
      mulM(srch, Mv);
      float q1 = 0.f, q2 = 0.f;
#pragma unroll
      for (int c = 0; c < ND; c++) { q1 = fmaf(srch[c], Ma[c] - fs[c], q1); q2 = fmaf(srch[c], Mv[c], q2); }
      const float jvl = rowdot(srch), tj = Dl * jvl;
      float alpha = 0.f, lo = 0.f, hi = -1.f;
      for (int ls = 0; ls < 40; ls++) {
        const float xx = fmaf(alpha, jvl, jr); const bool on = row && xx < 0.f;
        float d1 = on ? tj * xx : 0.f, d2 = on ? tj * jvl : 0.f, dn = on ? fabsf(tj * xx) : 0.f;
        if (ls == 0) warp_sum3(d1, d2, dn); else warp_sum2(d1, d2);
        d1 += q1 + alpha * q2; d2 += q2;
        if (ls == 0 && fabsf(d1) <= 2.4e-7f * (dn + fabsf(q1))) { alpha = 1.0f; break; }
        if (fabsf(d1) < 1e-6f * (1.0f + fabsf(q1))) break;
        if (d1 < 0.f) lo = alpha; else hi = alpha;
        if (!(d2 > 0.f)) break;
        float na = alpha - d1 / d2;
        if (hi > 0.f && (na <= lo || na >= hi)) na = 0.5f * (lo + hi);
        if (na < 0.f) na = 0.f;
        const bool done = fabsf(na - alpha) < 1e-7f * (1.0f + fabsf(alpha));
        alpha = na;
        if (done) break;
      }
      float amx = B2_NEWTON_AFLOOR;
#pragma unroll
      for (int c = 0; c < ND; c++) { a[c] = fmaf(alpha, srch[c], a[c]); Ma[c] = fmaf(alpha, Mv[c], Ma[c]); amx = fmaxf(amx, fmaxf(fabsf(a[c]), fabsf(as[c]))); }
      bool moving = false;
#pragma unroll
      for (int c = 0; c < ND; c++) moving |= c < nd && fabsf(alpha * srch[c]) > fmaf(B2_NEWTON_RTOL, amx, B2_NEWTON_NOISE * noise[c]);
      if (!moving) { it++; break; }
    }
    // efc_force at the solution (reporting only) and qfrc_constraint = M (a - a_smooth)
    const float sres = rowdot(a) - aref;
    if (row) { jar[lane] = sres < 0.f ? -Dl * sres : 0.f; Dr[lane] = Dl; }
    {
      float y[ND], My[ND];
#pragma unroll
      for (int c = 0; c < ND; c++) y[c] = a[c] - as[c];
      mulM(y, My);
#pragma unroll
      for (int c = 0; c < ND; c++) if (lane == c && c < nd) { const int d = cols.dof(c); p_qfc()[d] = My[c]; p_qacc()[d] = My[c]; }
    }
    sync();
    if (it >= iters && lane == 0 && counters) atomicAdd(&counters[CTR_ARENA_SPILL], 1ull);
    return it;
  }
  template <bool WD> __device__ __forceinline__ void solve_newton(unsigned long long* counters) {
    const int* madr = I(DI_dof_Madr); const int* ddepth = I(DI_dof_depth); const int* mcol = I(DI_Mcol);
    const int nisl = p_misc()[MISC_NISL], iters = dim(DD_iterations);
    const float scale = P.opt[DO_pgs_scale], tol = P.opt[DO_tolerance];
    int itmax = 0;
    for (int k = 0; k < nisl; k++) {
      const int n = p_isl_n()[k]; if (!n || p_isl_warp()[k] != wl) continue;
      const int nd = p_isl_nd()[k], ndp = r4(nd), ldj = p_isl_ldj()[k], e0 = p_isl_adr()[k], nh = nd * (nd + 1) / 2;
      if (W > 1 && nd > TEAMND) continue;      // solved by the whole team, below
      if (nd <= 6 && n <= 32) {                // a free body: in registers
#ifdef B2_PHASE_TIMING
        const long long ts0 = clock64();
#endif
        const int its = newton_island_small<WD>(k, counters);
#ifdef B2_PHASE_TIMING
        if (lane == 0 && B.phase_cycles) {
          atomicAdd(&B.phase_cycles[16], 1ull); atomicAdd(&B.phase_cycles[17], (unsigned long long)its);
          atomicAdd(&B.phase_cycles[18], (unsigned long long)n); atomicAdd(&B.phase_cycles[19], (unsigned long long)(clock64() - ts0));
        }
#endif
        itmax = max(itmax, its);
        continue;
      }
#ifdef B2_PHASE_TIMING
      const long long tn0 = clock64(); long long tq = tn0; const bool tbig = nd > 8;
      // islands above 8 dofs by phase: [22] setup, [23] J a + gradient, [24] H build, [25] + M and scaling, [26] Cholesky, [27] triangular solves,
      // [28] M s, J s, [29] line search, [30] update and exit tests, [31] finish
#define B2_NT(slot) do { if (tbig && lane == 0 && B.phase_cycles) { const long long t_ = clock64(); atomicAdd(&B.phase_cycles[slot], (unsigned long long)(t_ - tq)); tq = t_; } } while (0)
#else
#define B2_NT(slot) do { } while (0)
#endif
      const float* J = xs_J<WD>(k); const Cols cols = island_cols(k);       // wide: J (plain rows) and jv in the global workspace, H on chip
      float* H = island_A(k); float* a = H + r4(nh); float* as = a + ndp; float* fs = as + ndp;   // H: packed lower triangle
      float* Ma = fs + ndp; float* grad = Ma + ndp; float* srch = grad + ndp; float* Mv = srch + ndp; float* y = Mv + ndp;
      float* wrm = y + ndp; float* jv = WD ? xs_J<WD>(k) + r4(n * ldj) : wrm + ndp;
      float* Md = wrm + ndp + (WD ? 0 : r4(n));      // dense packed copy of the island's block of M
      float* Dr = xs_row_R<WD>() + e0; float* aref = xs_row_res<WD>() + e0; float* jar = xs_row_f<WD>() + e0; const float* bb = xs_row_b<WD>() + e0;
      for (int c = lane; c < nd; c += 32) { int d = cols.dof(c); as[c] = p_qas()[d]; fs[c] = p_qfs()[d]; wrm[c] = p_warm()[d]; }
      for (int q = lane; q < nh; q += 32) Md[q] = 0.f;
      sync();
      for (int c = lane; c < nd; c += 32) {
        int d = cols.dof(c), m0 = madr[d], dep = ddepth[d];
        for (int u = 0; u <= dep; u++) { int ca = p_dof_col()[mcol[m0 + u]]; int r = max(c, ca), cc = min(c, ca); Md[r * (r + 1) / 2 + cc] = p_M()[m0 + u]; }
      }
      // D = 1/R, capped at 1e8: MuJoCo floors R at 1e-15 (a body that cannot move along the row), and a penalty that
      // stiff puts J'D(Ja - aref) below fp32 resolution; the cap changes the constrained acceleration by <= force * 1e-8
      for (int i = lane; i < n; i += 32) {
        const float* Ji = J + i * ldj; float s = 0.f;
        for (int c = 0; c < nd; c++) s = fmaf(Ji[c], as[c], s);
        aref[i] = s - bb[i]; Dr[i] = fminf(1.0f / Dr[i], 1e8f);
      }
      // start from the cheaper of qacc_warmstart and qacc_smooth
      for (int c = lane; c < nd; c += 32) y[c] = wrm[c] - as[c];
      sync();
      warp_mulMd(Md, nd, y, Mv);
      float cw = 0.f, cs = 0.f;
      for (int i = lane; i < n; i += 32) {
        const float* Ji = J + i * ldj; float sw = 0.f, ss = 0.f;
        for (int c = 0; c < nd; c++) { sw = fmaf(Ji[c], wrm[c], sw); ss = fmaf(Ji[c], as[c], ss); }
        sw -= aref[i]; ss -= aref[i];
        if (sw < 0.f) cw += 0.5f * Dr[i] * sw * sw;
        if (ss < 0.f) cs += 0.5f * Dr[i] * ss * ss;
      }
      for (int c = lane; c < nd; c += 32) cw += 0.5f * Mv[c] * y[c];
      warp_sum2(cw, cs);
      for (int c = lane; c < nd; c += 32) a[c] = (cw < cs) ? wrm[c] : as[c];
      sync();
      warp_mulMd(Md, nd, a, Ma);
      B2_NT(22);
      int it = 0;
      for (; it < iters; it++) {
        for (int i = lane; i < n; i += 32) {
          const float* Ji = J + i * ldj; float s = 0.f;
          for (int c = 0; c < nd; c++) s = fmaf(Ji[c], a[c], s);
          s -= aref[i];
          jar[i] = s; jv[i] = s < 0.f ? Dr[i] : 0.f;      // D of the active rows, zero for the others (jv is free until the search direction exists)
        }
        sync();
        float g2 = 0.f; bool resolved = false;
        for (int c = lane; c < nd; c += 32) {
          float g = Ma[c] - fs[c], sabs = fabsf(Ma[c]) + fabsf(fs[c]);
          const float* Jc = J + c;
#pragma unroll 4
          for (int i = 0; i < n; i++) { const float t = Jc[i * ldj] * (jv[i] * jar[i]); g += t; sabs += fabsf(t); }      // no branch: the loads of the next rows are in flight
          grad[c] = g; g2 = fmaf(g, g, g2);
          resolved |= fabsf(g) > B2_NEWTON_GRADNOISE * sabs;
        }
        g2 = warp_sum(g2);
        B2_NT(23);
        if (scale * sqrtf(g2) < tol) break;
        // every gradient entry is below the rounding noise of its own terms: the Newton step it would produce is below the
        // step-size criterion at the end of the loop, so the factorisation is skipped (typically the second one of a solve)
        if (!__any_sync(B2_FULL, resolved)) break;
        // H = M + J' diag(D active) J, lower triangle, then symmetric diagonal scaling H <- S H S with S = diag(H_ii^-1/2):
        // joint inertias span orders of magnitude (finger hinges vs the crane), which fp32 Cholesky does not survive
        // unscaled; the scaled matrix has a unit diagonal
        if (nd >= 16 && nd <= 128 && (nd <= 64 || B2_NEWTON_SPARSE128)) {
          // J rows are sparse (a contact touches the dof chains of its two bodies only): per group of rows that belong
          // to one contact / one limit, list the columns any active row touches and accumulate only their pairs (pays
          // from ~16 dofs; a free body's 6-dof island has dense rows and takes the plain loop below).  Up to 128 columns:
          // construction's humanoid standing in the materials is one island of 60-99 dofs, where the plain loop costs
          // nd^2 n / 64 multiply-adds per lane
          for (int q = lane; q < nh; q += 32) H[q] = 0.f;
          sync();
          const int* rinfo = xs_row_info<WD>() + e0; int* clist = reinterpret_cast<int*>(y);
          const unsigned lt = (1u << lane) - 1u;
          for (int i0 = 0; i0 < n;) {
            int g = 1; const int key = rinfo[i0] >> 4;
            while (i0 + g < n && (rinfo[i0 + g] >> 4) == key) g++;
            unsigned act = 0;
            for (int r = 0; r < g; r++) if (jar[i0 + r] < 0.f) act |= 1u << r;
            if (act) {
              bool nz[4] = {false, false, false, false};
              for (int r = 0; r < g; r++) if ((act >> r) & 1) {
                const float* Jr = J + (i0 + r) * ldj;
#pragma unroll
                for (int u = 0; u < 4; u++) nz[u] |= lane + 32 * u < nd && Jr[lane + 32 * u] != 0.f;
              }
              int kk = 0;
#pragma unroll
              for (int u = 0; u < 4; u++) {
                if (32 * u < nd) {
                  const unsigned m = __ballot_sync(B2_FULL, nz[u]);
                  if (nz[u]) clist[kk + __popc(m & lt)] = lane + 32 * u;
                  kk += __popc(m);
                }
              }
              sync();
              for (int q = lane; q < kk * (kk + 1) / 2; q += 32) {
                int a = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
                while (a * (a + 1) / 2 > q) a--;
                while ((a + 1) * (a + 2) / 2 <= q) a++;
                const int ca = clist[a], cb = clist[q - a * (a + 1) / 2]; float h = 0.f;
                for (int r = 0; r < g; r++) if ((act >> r) & 1) { const float* Jr = J + (i0 + r) * ldj; h = fmaf(Jr[ca] * Dr[i0 + r], Jr[cb], h); }
                H[ca * (ca + 1) / 2 + cb] += h;
              }
              sync();
            }
            i0 += g;
          }
        } else {
          for (int q = lane; q < nh; q += 32) {
            int r = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
            while (r * (r + 1) / 2 > q) r--;
            while ((r + 1) * (r + 2) / 2 <= q) r++;
            int c = q - r * (r + 1) / 2; float h = 0.f;
            const float* Jr = J + r; const float* Jc = J + c;
#pragma unroll 4
            for (int i = 0; i < n; i++) h = fmaf(Jr[i * ldj] * jv[i], Jc[i * ldj], h);
            H[q] = h;
          }
          sync();
        }
        B2_NT(24);
        for (int q = lane; q < nh; q += 32) H[q] += Md[q];
        sync();
        for (int c = lane; c < nd; c += 32) Mv[c] = rsqrtf(fmaxf(H[c * (c + 1) / 2 + c], 1e-30f));   // Mv doubles as the scale vector here
        sync();
        for (int r = 0; r < nd; r++) { float sr = Mv[r]; float* Hr = H + r * (r + 1) / 2; for (int c = lane; c <= r; c += 32) Hr[c] *= sr * Mv[c]; }
        sync();
        B2_NT(25);
        // Cholesky H = L L' in place (right-looking; lanes over the rows below the pivot).  The reciprocal pivots go to wrm
        // (the warm start is not needed once the iteration has begun), so a pivot costs one rsqrt and two warp barriers and
        // the triangular solves below multiply instead of dividing
        for (int p = 0; p < nd; p++) {
          sync();
          const float inv = rsqrtf(fmaxf(H[p * (p + 1) / 2 + p], 1e-7f));
          if (lane == 0) wrm[p] = inv;
          for (int i = p + 1 + lane; i < nd; i += 32) H[i * (i + 1) / 2 + p] *= inv;
          sync();
          for (int i = p + 1 + lane; i < nd; i += 32) {
            float* Hi = H + i * (i + 1) / 2; float lip = Hi[p];
#pragma unroll 4
            for (int c = p + 1; c <= i; c++) Hi[c] = fmaf(-lip, H[c * (c + 1) / 2 + p], Hi[c]);
          }
        }
        B2_NT(26);
        // search = -H^-1 grad: L z = -S grad, L' x = z, both column-oriented (once an unknown is final every lane subtracts
        // its multiple from the entries it owns): one warp barrier per pivot and no reductions.  z lives in srch, x in y.
        for (int c = lane; c < nd; c += 32) y[c] = -grad[c] * Mv[c];
        for (int r = 0; r < nd; r++) {
          sync();
          const float zr = y[r] * wrm[r];
          if (lane == 0) srch[r] = zr;
          for (int i = r + 1 + lane; i < nd; i += 32) y[i] = fmaf(-H[i * (i + 1) / 2 + r], zr, y[i]);
        }
        for (int r = nd - 1; r >= 0; r--) {
          sync();
          const float* Hr = H + r * (r + 1) / 2;
          const float xr = srch[r] * wrm[r];
          if (lane == 0) y[r] = xr;
          for (int i = lane; i < r; i += 32) srch[i] = fmaf(-Hr[i], xr, srch[i]);
        }
        sync();
        B2_NT(27);
        // undo the scaling, s = S (S H S)^-1 S (-g); y <- fp32 noise floor of each acceleration, (|f_smooth| + |M a|) / H_cc
        for (int c = lane; c < nd; c += 32) { float sc = Mv[c]; srch[c] = y[c] * sc; y[c] = sc * sc * (fabsf(fs[c]) + fabsf(Ma[c])); }
        sync();
        warp_mulMd(Md, nd, srch, Mv);
        float q1 = 0.f, q2 = 0.f;
        for (int c = lane; c < nd; c += 32) { q1 = fmaf(srch[c], Ma[c] - fs[c], q1); q2 = fmaf(srch[c], Mv[c], q2); }
        for (int i = lane; i < n; i += 32) {
          const float* Ji = J + i * ldj; float sj = 0.f;
          for (int c = 0; c < nd; c++) sj = fmaf(Ji[c], srch[c], sj);
          jv[i] = sj;
        }
        warp_sum2(q1, q2);
        sync();
        B2_NT(28);
        // exact minimisation along the search direction
        float alpha = 0.f, lo = 0.f, hi = -1.f;
        for (int ls = 0; ls < 40; ls++) {
          float d1 = 0.f, d2 = 0.f, dn = 0.f;
          for (int i = lane; i < n; i += 32) { float x = fmaf(alpha, jv[i], jar[i]); if (x < 0.f) { float t = Dr[i] * jv[i]; d1 = fmaf(t, x, d1); d2 = fmaf(t, jv[i], d2); dn += fabsf(t * x); } }
          if (ls == 0) warp_sum3(d1, d2, dn); else warp_sum2(d1, d2);
          d1 += q1 + alpha * q2; d2 += q2;
          // a slope at alpha = 0 that is below the rounding noise of its own terms carries no information: take the plain
          // Newton step (the direction is noise-sized too) instead of searching on noise
          if (ls == 0 && fabsf(d1) <= 2.4e-7f * (dn + fabsf(q1))) { alpha = 1.0f; break; }
          if (fabsf(d1) < 1e-6f * (1.0f + fabsf(q1))) break;
          if (d1 < 0.f) lo = alpha; else hi = alpha;
          if (!(d2 > 0.f)) break;
          float na = alpha - d1 / d2;
          if (hi > 0.f && (na <= lo || na >= hi)) na = 0.5f * (lo + hi);
          if (na < 0.f) na = 0.f;
          bool done = fabsf(na - alpha) < 1e-7f * (1.0f + fabsf(alpha));
          alpha = na;
          if (done) break;
        }
        B2_NT(29);
        // fp32 termination: no acceleration moves by more than 1e-5 of the island's largest one plus its own rounding
        // floor (the gradient test above cannot fire once the stiff rows' rounding noise exceeds the tolerance)
        float amx = B2_NEWTON_AFLOOR;
        for (int c = lane; c < nd; c += 32) { a[c] = fmaf(alpha, srch[c], a[c]); Ma[c] = fmaf(alpha, Mv[c], Ma[c]); amx = fmaxf(amx, fmaxf(fabsf(a[c]), fabsf(as[c]))); }
#pragma unroll
        for (int o = 16; o; o >>= 1) amx = fmaxf(amx, __shfl_xor_sync(0xffffffffu, amx, o));
        bool moving = false;
        for (int c = lane; c < nd; c += 32) moving |= fabsf(alpha * srch[c]) > fmaf(B2_NEWTON_RTOL, amx, B2_NEWTON_NOISE * y[c]);
        moving = __any_sync(0xffffffffu, moving);
        sync();
#ifdef B2_NEWTON_DEBUG
        if (it >= iters - 6 && lane == 0) printf("newton blk %d wb %d isl %d nd %d n %d it %d |g| %.4g alpha %.6g amx %.4g q1 %.4g q2 %.4g\n", (int)blockIdx.x, wb, k, nd, n, it, sqrtf(g2), alpha, amx, q1, q2);
#endif
        B2_NT(30);
        if (!moving) { it++; break; }
      }
      // efc_force at the solution (reporting only: a stiff row's force is below fp32 resolution of J a - aref).  The
      // constraint force that drives the step comes from the optimality condition instead, qfrc_constraint =
      // M a - qfrc_smooth = M (a - a_smooth), which is accurate to fp32 in a whatever the row stiffness.
      for (int i = lane; i < n; i += 32) {
        const float* Ji = J + i * ldj; float s = 0.f;
        for (int c = 0; c < nd; c++) s = fmaf(Ji[c], a[c], s);
        s -= aref[i];
        jar[i] = s < 0.f ? -Dr[i] * s : 0.f;
      }
      for (int c = lane; c < nd; c += 32) y[c] = a[c] - as[c];
      sync();
      warp_mulMd(Md, nd, y, Mv);
      for (int c = lane; c < nd; c += 32) { int d = cols.dof(c); p_qfc()[d] = Mv[c]; p_qacc()[d] = Mv[c]; }
      sync();
      if (it >= iters && lane == 0 && counters) atomicAdd(&counters[CTR_ARENA_SPILL], 1ull);
#ifdef B2_PHASE_TIMING
      B2_NT(31);
      if (lane == 0 && B.phase_cycles) {      // [16..19] islands <= 8 dofs: solves, iterations, rows, cycles; [20] [21] larger islands: solves, iterations
        if (!tbig) {
          atomicAdd(&B.phase_cycles[16], 1ull); atomicAdd(&B.phase_cycles[17], (unsigned long long)it);
          atomicAdd(&B.phase_cycles[18], (unsigned long long)n); atomicAdd(&B.phase_cycles[19], (unsigned long long)(clock64() - tn0));
        } else { atomicAdd(&B.phase_cycles[20], 1ull); atomicAdd(&B.phase_cycles[21], (unsigned long long)it); }
      }
#endif
      itmax = max(itmax, it);
    }
    if (W > 1) {
      for (int k = 0; k < nisl; k++)
        if (p_isl_n()[k] && p_isl_nd()[k] > TEAMND) itmax = max(itmax, newton_island_team<WD>(k, counters));
      team_sync();
    }
    if (lane == 0) p_red()[wl] = (float)itmax;
    team_sync();
    if (tl == 0) {
      int itall = 0;
      for (int q = 0; q < W; q++) itall = max(itall, (int)p_red()[q]);
      p_misc()[MISC_ITERS] = itall; if (counters) atomicAdd(&counters[CTR_SOLVER_ITERS], (unsigned long long)itall);
    }
    team_sync();
  }

  // ---- qfrc_constraint = J' f, also copied into qacc as the right-hand side of the pass-1 solve
  template <bool WD> __device__ __forceinline__ void qfrc_constraint(bool newton) {
    const int* dtree = I(DI_dof_tree);
    int nv = dim(DD_nv);
#pragma unroll 1
    for (int d = lane; d < nv; d += 32) {
      int k = p_tree_isl()[dtree[d]]; int n = p_isl_n()[k]; float s0 = 0.f, s1 = 0.f;
      if (newton) { if (n && p_misc()[MISC_NEFC] > 0) continue; }
      else if (n && p_misc()[MISC_NEFC] > 0) {
        int ldj = p_isl_ldj()[k], e0 = p_isl_adr()[k], c = p_dof_col()[d]; const float* J = xs_J<WD>(k) + c; const float* f = xs_row_f<WD>() + e0;
        const int blk = jblk(ldj);
        int i = 0;
        for (; i + 2 <= n; i += 2) { s0 = fmaf(J[jr<WD>(i, blk, ldj)], f[i], s0); s1 = fmaf(J[jr<WD>(i + 1, blk, ldj)], f[i + 1], s1); }
        if (i < n) s0 = fmaf(J[jr<WD>(i, blk, ldj)], f[i], s0);
      }
      p_qfc()[d] = s0 + s1; p_qacc()[d] = s0 + s1;
    }
    sync();
  }

#ifdef B2_PHASE_TIMING
#define B2_TICK(k) do { long long t_ = clock64(); if (tl == 0 && B.phase_cycles) atomicAdd(&B.phase_cycles[k], (unsigned long long)(t_ - tphase)); tphase = t_; } while (0)
#else
#define B2_TICK(k) do { } while (0)
#endif
  __device__ void reset_data() {   // warp-level (called by warp 0)
    const float* q0 = F(DF_qpos0);
    int nq = dim(DD_nq), nv = dim(DD_nv), nu = dim(DD_nu);
    for (int i = lane; i < nq; i += 32) p_qpos()[i] = q0[i];
    for (int i = lane; i < nv; i += 32) { p_qvel()[i] = 0.f; p_warm()[i] = 0.f; p_qapp()[i] = 0.f; }
    for (int i = lane; i < nu; i += 32) p_ctrl()[i] = 0.f;
    sync();
  }
  __device__ bool bad_state(const float* x, int n) {
    int bad = 0;
    for (int i = lane; i < n; i += 32) { float v = x[i]; if (!(v == v) || fabsf(v) > B2_MAXVAL) bad = 1; }
    return __any_sync(B2_FULL, bad);
  }

  // ---- q_to = q_from (+) h * vel  (mj_integratePos: free-joint quaternions by the exponential map)
  __device__ __forceinline__ void integrate_pos(float* q_to, const float* q_from, const float* vel, float h) {
    const int* jtype = I(DI_jnt_type); const int* jq = I(DI_jnt_qposadr); const int* jd = I(DI_jnt_dofadr);
    int njnt = dim(DD_njnt);
#pragma unroll 1
    for (int j = lane; j < njnt; j += 32) {
      int qa = jq[j], da = jd[j];
      if (jtype[j] == 0) {
        for (int k = 0; k < 3; k++) q_to[qa + k] = q_from[qa + k] + h * vel[da + k];
        V3 om = ld3(vel + da + 3); float n = norm(om);
        Q4 q = ldq(q_from + qa + 3);
        if (n >= B2_MINVAL) q = qnormalize(qmul(q, axisangle(om * (1.f / n), h * n)));
        stq(q_to + qa + 3, q);
      } else q_to[qa] = q_from[qa] + h * vel[da];
    }
  }

  // ---- mj_step -- SURVEY B.0 / B.7 (integrate == false: mj_forward).
  // One forward evaluation is a loop over passes that share ONE call site of factor() and solve():
  //   pass 0  velocities, CRB, M, bias | contacts, rows     -> factor(M)        -> qacc_smooth = M^-1 qfrc_smooth
  //   pass 1  J, A, PGS (whole team), qfrc_constraint       ->                    qacc = qacc_smooth + M^-1 qfrc_constraint
  //   pass 2  (mj_Euler only)                               -> factor(M + h D)  -> qacc' = (M + h D)^-1 (qfrc_smooth + qfrc_constraint)
  // Warp 0 runs the dynamics chain while warp 1 % W runs the contact chain; the NaN retry of mj_step (mj_checkAcc) is the
  // attempt loop.  Euler (implicit joint damping) is one evaluation; RK4 (mj_RungeKutta, N = 4) is four evaluations
  // around the same loop body: the classical tableau is diagonal, so only q0, v0 and the two running weighted sums of
  // stage velocities / accelerations are kept.  qacc_warmstart is refreshed once per mj_step, from the last stage.
  __device__ __forceinline__ void step(unsigned long long* counters, bool integrate) {
    float* time = p_time();
    int nq = dim(DD_nq), nv = dim(DD_nv); float h = P.opt[DO_timestep];
    const bool rk4 = dim(DD_integrator) == 1;
    const int nstage = (integrate && rk4) ? 4 : 1;
    const int npass = (integrate && !rk4) ? 3 : 2;
#ifdef B2_PHASE_TIMING
    long long tphase = clock64();
#endif
    if (integrate && wl == 0) {
      if (bad_state(p_qpos(), nq) | bad_state(p_qvel(), nv)) {
        reset_data(); if (lane == 0) { *time = 0.f; if (counters) atomicAdd(&counters[CTR_NAN_RESET], 1ull); }
      }
    }
#pragma unroll 1
    for (int stage = 0; stage < nstage; stage++) {
#pragma unroll 1
      for (int attempt = 0; attempt < 2; attempt++) {
        if (B.lockstep) asm volatile("bar.sync 0;" ::: "memory");      // every thread of the CTA: teams without work keep arriving (b2_env_kernel)
        if (wl == 0) { kinematics(); com_pos(); }
        team_sync(); B2_TICK(0);
        bool restart = false;
#pragma unroll 1
        for (int pass = 0; pass < npass; pass++) {
          if (pass == 0) {
            if (wl == 0) { vel_pass(); backward_pass(); mass_and_smooth(); }
#ifdef B2_PHASE_TIMING
            if (wl == (1 % W)) {      // the contact chain's own clock: [8] collision, [6] make_rows
              const long long c0 = clock64(); collision(counters); const long long c1 = clock64(); make_rows(counters); const long long c2 = clock64();
              if (lane == 0 && B.phase_cycles) { atomicAdd(&B.phase_cycles[8], (unsigned long long)(c1 - c0)); atomicAdd(&B.phase_cycles[6], (unsigned long long)(c2 - c1)); }
            }
#else
            if (wl == (1 % W)) { collision(counters); make_rows(counters); }
#endif
            if (W == 3 && npass == 3) {
              // mj_Euler's factor of M + h D does not depend on the constraint solve: the team's third warp builds it
              // as soon as M exists (a two-warp named barrier hands M over), beside warp 0's factor of M
              if (wl == 0 || wl == 2) asm volatile("bar.sync %0, 64;" ::"r"(barid + 6) : "memory");
              if (wl == 2) factor(h, true);
            }
          } else if (pass == 1) {
            if (wide) { wide_pass(counters); B2_TICK(12); }      // cold path, out of line
            else {
              if (p_misc()[MISC_NEFC] > 0) {
                B2_TICK(7); fill_rows<false>(); B2_TICK(9);
                if (newton()) solve_newton<false>(counters);
                else { build_A(); B2_TICK(10); solve_pgs(counters); }
                B2_TICK(11);
              }
              else if (tl == 0) p_misc()[MISC_ITERS] = 0;
              if (wl == 0) qfrc_constraint<false>(newton());
            }
          } else {
            if (wl == 0) { for (int d = lane; d < nv; d += 32) p_tmp()[d] = p_qfs()[d] + p_qfc()[d]; sync(); }
          }
          if (wl == 0) {
            if (pass == 0) factor(0.f, false);
            else if (pass == 2 && W != 3) factor(h, true);
            float* x = b2_smem + wb + (pass == 0 ? B.off.qas : pass == 1 ? B.off.qacc : B.off.tmp);
            if (pass == 0) { for (int d = lane; d < nv; d += 32) x[d] = p_qfs()[d]; sync(); }
            solve(x, pass == 2);
            if (pass == 1) { for (int d = lane; d < nv; d += 32) x[d] += p_qas()[d]; sync(); }
          }
          team_sync(); B2_TICK(1 + pass);
          if (pass == 0) refresh_wide();
          if (pass == 1 && integrate && stage == 0 && attempt == 0) {      // mj_checkAcc
            if (wl == 0) { bool bad = bad_state(p_qacc(), nv); if (lane == 0) p_misc()[MISC_FLAG] = bad ? 1 : 0; }
            team_sync();
            if (p_misc()[MISC_FLAG]) {
              if (wl == 0) { reset_data(); if (lane == 0) { *time = 0.f; if (counters) atomicAdd(&counters[CTR_NAN_RESET], 1ull); } }
              team_sync(); restart = true; break;
            }
          }
        }
        if (!restart) break;
      }
      // mj_fwdConstraint ends by saving qacc into qacc_warmstart (qacc_smooth when there are no rows: qacc equals it then), so
      // RK4 stages 2-4 start from the previous stage's solution and mj_forward moves the warm start as well
      if (!B.warm_once && wl == 0) {
#pragma unroll 1
        for (int d = lane; d < nv; d += 32) p_warm()[d] = p_qacc()[d];
        sync();
      }
      if (nstage == 4) {
        if (wl == 0) {
          const float bw = (stage == 0 || stage == 3) ? (1.0f / 6.0f) : (1.0f / 3.0f);
          const float cs = stage == 2 ? 1.0f : 0.5f;
          if (stage == 0) {
#pragma unroll 1
            for (int i = lane; i < nq; i += 32) p_rk_q0()[i] = p_qpos()[i];
          }
#pragma unroll 1
          for (int d = lane; d < nv; d += 32) {
            float v = p_qvel()[d], a = p_qacc()[d];
            if (stage == 0) { p_rk_v0()[d] = v; p_rk_sv()[d] = bw * v; p_rk_sa()[d] = bw * a; }
            else { p_rk_sv()[d] = fmaf(bw, v, p_rk_sv()[d]); p_rk_sa()[d] = fmaf(bw, a, p_rk_sa()[d]); }
          }
          sync();
          if (stage < 3) {
            integrate_pos(p_qpos(), p_rk_q0(), p_qvel(), h * cs);     // uses this stage's velocity
            sync();
#pragma unroll 1
            for (int d = lane; d < nv; d += 32) p_qvel()[d] = fmaf(h * cs, p_qacc()[d], p_rk_v0()[d]);
          } else {
            integrate_pos(p_qpos(), p_rk_q0(), p_rk_sv(), h);
#pragma unroll 1
            for (int d = lane; d < nv; d += 32) { p_qvel()[d] = fmaf(h, p_rk_sa()[d], p_rk_v0()[d]); p_warm()[d] = p_qacc()[d]; }
            if (lane == 0) { *time += h; if (counters) atomicAdd(&counters[CTR_SUBSTEPS], 1ull); }
          }
          sync();
        }
        team_sync();
      }
    }
    if (!integrate || rk4) { B2_TICK(13); return; }
    if (wl == 0) {
#pragma unroll 1
      for (int d = lane; d < nv; d += 32) { p_qvel()[d] += h * p_tmp()[d]; p_warm()[d] = p_qacc()[d]; }
      sync();
      integrate_pos(p_qpos(), p_qpos(), p_qvel(), h);
      if (lane == 0) { *time += h; if (counters) atomicAdd(&counters[CTR_SUBSTEPS], 1ull); }
    }
    team_sync();
    B2_TICK(13);
  }
};

// ---- stage the model tables into shared memory: one TMA bulk copy per buffer, completion on an mbarrier
__device__ inline void stage_model(const DevModel& P, int* smi, float* smf, uint64_t* bar) {
  uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t bytes_i = (uint32_t)P.n_ints_staged * 4u, bytes_f = (uint32_t)P.n_flts * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes_i + bytes_f) : "memory");
    uint32_t dst_i = (uint32_t)__cvta_generic_to_shared(smi), dst_f = (uint32_t)__cvta_generic_to_shared(smf);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_i),
                 "l"(P.ints), "r"(bytes_i), "r"(bar_s) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_f),
                 "l"(P.flts), "r"(bytes_f), "r"(bar_s) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar_s), "r"(0u) : "memory");
  }
}

}  // namespace b2
