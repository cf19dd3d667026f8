// b2_engine.cuh -- lock-step mj_step for one environment per CTA (T = 32*W threads), sm_100a.
//
// Replaces the arithmetic behind the reference's `mujoco.mj_step(model, data)` call sites
// (quadruped_parkour_env/parkour_env.py:348,368 and siblings) for the feature
// subset in SURVEY App. B.  Design:
//   * model tables are staged once per CTA into shared memory with one TMA bulk copy per buffer
//     (cp.async.bulk + mbarrier);
//   * every per-env intermediate (frames, spatial inertias, sparse M and its L'DL factor, contacts, J, A) lives in
//     shared memory for the whole env-step, HBM is touched only to load/store qpos/qvel/warmstart/ctrl/task state;
//   * tree passes are level-synchronous over bodies / dofs; the constraint problem is split into islands
//     (kinematic trees that can exchange contact forces); each island's A = J M^-1 J' + R is dense in shared
//     memory and its projected Gauss-Seidel sweep is owned by one warp with forces and residuals held in
//     registers (one shuffle broadcast + S FMAs per row).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2_device_layout.h"
#include "b2_math.cuh"

namespace b2 {

#define B2_MINVAL 1e-15f
#define B2_MAXVAL 1e10f
#define B2_MINIMP 0.0001f
#define B2_MAXIMP 0.9999f
#define B2_MAX_ISLANDS 16
#define B2_CON_STRIDE 16   // floats per contact record
#define B2_PGS_S 3         // 32*S rows per island held in registers

struct DevModel {
  const int* ints; const float* flts;
  int n_ints, n_flts;
  int ioff[B2DEV_N_INT_FIELDS];
  int foff[B2DEV_N_FLT_FIELDS];
  int dim[DD_COUNT];
  float opt[DO_COUNT];
};

enum { CTR_NAN_RESET = 0, CTR_CON_DROPPED = 1, CTR_ROW_DROPPED = 2, CTR_ARENA_OVERFLOW = 3, CTR_EPISODES = 4,
       CTR_SOLVER_ITERS = 5, CTR_SUBSTEPS = 6, CTR_COUNT = 8 };

struct BatchView {
  int n_envs;
  float *qpos, *qvel, *warm, *ctrl, *qfrc_applied, *time;
  int nqp, nvp, nup;               // row pitches (floats)
  int *ti; float *tf; int nti, ntf;  // per-env task state
  float *obs, *final_obs, *reward; uint8_t *term, *trunc; int obs_dim;
  const float* action; int act_dim;
  const uint8_t* reset_mask;
  int *c_ncon, *c_geom; float* c_dist; int c_cap;   // optional contact export
  float* xpos_out;                                   // optional [N][nbody*3] export of the last forward pass
  float* debug_out; int debug_n;                     // optional [N][debug_n] dump of solver intermediates (bring-up / tests)
  unsigned long long* counters;                      // [N][CTR_COUNT]
  unsigned long long seed;
  int arena_floats, con_cap, row_cap;
  int nsub;                                          // physics sub-steps for MODE_PHYS
  int env_offset;                                    // global index of env 0 (multi-GPU sharding keeps RNG streams fixed)
};

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_PHYS = 2, MODE_FORWARD = 3 };

// -------------------------------------------------------------------------------------------- workspace
struct Ws {
  // model (shared-memory copies)
  const int* mi; const float* mf;
  // state
  float *qpos, *qvel, *warm, *ctrl, *qapp;
  // kinematics / dynamics
  float *xpos, *xquat, *xmat, *xipos, *cinert, *crb, *cdof, *cvel, *cacc, *cfrc, *rootcom;
  float *M, *LD, *invD, *qfs, *qas, *qfc, *qacc, *tmp;
  // contacts and rows
  float* con; int* lim_row; int* con_row;
  int* row_info; float *row_pos, *row_margin, *row_R, *row_aref, *row_b, *row_f, *row_D;
  int *isl_n, *isl_adr, *isl_J, *isl_A, *isl_ldj;
  float* arena; float* red; int* misc; float* time;
};
enum { MISC_NCON = 0, MISC_NEFC = 1, MISC_FLAG = 2, MISC_ITERS = 3, MISC_ARENA_USED = 4, MISC_COUNT = 8 };

__host__ __device__ inline int r4(int n) { return (n + 3) & ~3; }

// Shared-memory carve-up; the host calls this with ws==nullptr to size the launch.
__host__ __device__ inline size_t ws_layout(const int* dim, int con_cap, int row_cap, int arena_floats, int n_ints,
                                            int n_flts, Ws* ws, unsigned char* base) {
  size_t off = 0;
  auto take = [&](size_t n_words) { size_t o = off; off += (size_t)r4((int)n_words) * 4; return base ? base + o : (unsigned char*)0; };
  int nq = dim[DD_nq], nv = dim[DD_nv], nu = dim[DD_nu], nb = dim[DD_nbody], nM = dim[DD_nM];
  int nroot = dim[DD_nroot], nlim = dim[DD_nlim];
  unsigned char* p;
  p = take(n_ints); if (ws) ws->mi = (const int*)p;
  p = take(n_flts); if (ws) ws->mf = (const float*)p;
#define TAKEF(name, n) p = take(n); if (ws) ws->name = (float*)p;
#define TAKEI(name, n) p = take(n); if (ws) ws->name = (int*)p;
  TAKEF(qpos, nq) TAKEF(qvel, nv) TAKEF(warm, nv) TAKEF(ctrl, nu > 0 ? nu : 1) TAKEF(qapp, nv)
  TAKEF(xpos, 3 * nb) TAKEF(xquat, 4 * nb) TAKEF(xmat, 9 * nb) TAKEF(xipos, 3 * nb)
  TAKEF(cinert, 10 * nb) TAKEF(crb, 10 * nb) TAKEF(cdof, 6 * nv) TAKEF(cvel, 6 * nb) TAKEF(cacc, 6 * nb)
  TAKEF(cfrc, 6 * nb) TAKEF(rootcom, 3 * (nroot > 0 ? nroot : 1))
  TAKEF(M, nM) TAKEF(LD, nM) TAKEF(invD, nv) TAKEF(qfs, nv) TAKEF(qas, nv) TAKEF(qfc, nv) TAKEF(qacc, nv) TAKEF(tmp, nv)
  TAKEF(con, con_cap * B2_CON_STRIDE) TAKEI(lim_row, 2 * (nlim > 0 ? nlim : 1)) TAKEI(con_row, con_cap)
  TAKEI(row_info, row_cap) TAKEF(row_pos, row_cap) TAKEF(row_margin, row_cap) TAKEF(row_R, row_cap)
  TAKEF(row_aref, row_cap) TAKEF(row_b, row_cap) TAKEF(row_f, row_cap) TAKEF(row_D, row_cap)
  TAKEI(isl_n, B2_MAX_ISLANDS) TAKEI(isl_adr, B2_MAX_ISLANDS + 1) TAKEI(isl_J, B2_MAX_ISLANDS)
  TAKEI(isl_A, B2_MAX_ISLANDS) TAKEI(isl_ldj, B2_MAX_ISLANDS)
  TAKEF(arena, arena_floats) TAKEF(red, 64) TAKEI(misc, MISC_COUNT) TAKEF(time, 4)
#undef TAKEF
#undef TAKEI
  return off + 16;  // + mbarrier
}

// -------------------------------------------------------------------------------------------- engine
template <int T>
struct Engine {
  static constexpr int W = T / 32;
  const DevModel& P;
  Ws w;
  int tid, lane, warp;

  __device__ Engine(const DevModel& p) : P(p) { tid = threadIdx.x; lane = tid & 31; warp = tid >> 5; }

  __device__ __forceinline__ const int* I(int f) const { return w.mi + P.ioff[f]; }
  __device__ __forceinline__ const float* F(int f) const { return w.mf + P.foff[f]; }
  __device__ __forceinline__ int dim(int k) const { return P.dim[k]; }
  __device__ __forceinline__ void sync() const { if (T == 32) __syncwarp(); else __syncthreads(); }

  // ---- stage the model tables into shared memory: one TMA bulk copy per buffer, completion on an mbarrier
  __device__ void stage_model(unsigned char* smem, size_t bar_off) {
    uint64_t* bar = (uint64_t*)(smem + bar_off);
    uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t bytes_i = (uint32_t)P.n_ints * 4u, bytes_f = (uint32_t)P.n_flts * 4u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes_i + bytes_f) : "memory");
      uint32_t dst_i = (uint32_t)__cvta_generic_to_shared(w.mi), dst_f = (uint32_t)__cvta_generic_to_shared(w.mf);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_i),
                   "l"(P.ints), "r"(bytes_i), "r"(bar_s)
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_f),
                   "l"(P.flts), "r"(bytes_f), "r"(bar_s)
                   : "memory");
    }
    // all threads wait for phase 0
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(bar_s), "r"(0u) : "memory");
    }
  }

  // ---- B.1 kinematics: level-synchronous over bodies
  __device__ void kinematics() {
    const int* parent = I(DI_body_parentid); const int* jadr = I(DI_body_jntadr); const int* jnum = I(DI_body_jntnum);
    const int* jtype = I(DI_jnt_type); const int* jq = I(DI_jnt_qposadr); const int* jd = I(DI_jnt_dofadr);
    const int* ladr = I(DI_level_adr); const int* lbody = I(DI_level_body);
    const float* bpos = F(DF_body_pos); const float* bquat = F(DF_body_quat); const float* bipos = F(DF_body_ipos);
    const float* jpos = F(DF_jnt_pos); const float* jaxis = F(DF_jnt_axis); const float* q0 = F(DF_qpos0);
    if (tid == 0) {
      st3(w.xpos, v3(0, 0, 0)); Q4 qi; qi.w = 1; qi.x = qi.y = qi.z = 0; stq(w.xquat, qi); quat2mat(w.xmat, qi);
      st3(w.xipos, v3(0, 0, 0));
      for (int k = 0; k < 6; k++) { w.cvel[k] = 0.f; w.cfrc[k] = 0.f; }
      for (int k = 0; k < 10; k++) { w.cinert[k] = 0.f; w.crb[k] = 0.f; }
      w.cacc[0] = w.cacc[1] = w.cacc[2] = 0.f;
      w.cacc[3] = -P.opt[DO_gx]; w.cacc[4] = -P.opt[DO_gy]; w.cacc[5] = -P.opt[DO_gz];
    }
    sync();
    int maxdepth = dim(DD_maxdepth);
    for (int l = 1; l <= maxdepth; l++) {
      for (int idx = ladr[l] + tid; idx < ladr[l + 1]; idx += T) {
        int b = lbody[idx], p = parent[b], ja = jadr[b], jn = jnum[b];
        V3 pos; Q4 quat;
        if (jn == 1 && jtype[ja] == 0) {
          float* q = w.qpos + jq[ja];
          pos = ld3(q); quat = qnormalize(ldq(q + 3)); stq(q + 3, quat);
          int d = jd[ja];
          float mat[9]; quat2mat(mat, quat);
#pragma unroll
          for (int a = 0; a < 3; a++) {
            V3 e = v3(a == 0, a == 1, a == 2);
            st3(w.cdof + 6 * (d + a), v3(0, 0, 0)); st3(w.cdof + 6 * (d + a) + 3, e);
            V3 ax = matcol(mat, a);
            st3(w.cdof + 6 * (d + 3 + a), ax); st3(w.cdof + 6 * (d + 3 + a) + 3, cross(ax, pos) * -1.f);
          }
        } else {
          pos = ld3(w.xpos + 3 * p) + mulmat(w.xmat + 9 * p, ld3(bpos + 3 * b));
          quat = qmul(ldq(w.xquat + 4 * p), ldq(bquat + 4 * b));
          for (int k = 0; k < jn; k++) {
            int j = ja + k, d = jd[j];
            V3 lp = ld3(jpos + 3 * j), la = ld3(jaxis + 3 * j);
            V3 anchor = qrot(quat, lp) + pos, axis = qrot(quat, la);
            float dq = w.qpos[jq[j]] - q0[jq[j]];
            if (jtype[j] == 2) {  // slide
              pos = pos + axis * dq;
              st3(w.cdof + 6 * d, v3(0, 0, 0)); st3(w.cdof + 6 * d + 3, axis);
            } else {              // hinge
              quat = qmul(quat, axisangle(la, dq));
              pos = anchor - qrot(quat, lp);
              st3(w.cdof + 6 * d, axis); st3(w.cdof + 6 * d + 3, cross(axis, anchor) * -1.f);  // + axis x com later
            }
          }
          quat = qnormalize(quat);
        }
        st3(w.xpos + 3 * b, pos); stq(w.xquat + 4 * b, quat);
        float mat[9]; quat2mat(mat, quat);
#pragma unroll
        for (int k = 0; k < 9; k++) w.xmat[9 * b + k] = mat[k];
        st3(w.xipos + 3 * b, pos + mulmat(mat, ld3(bipos + 3 * b)));
      }
      sync();
    }
  }

  // ---- mj_comPos: com of each root subtree, cinert about it, cdof offset fix-up
  __device__ void com_pos() {
    const int* radr = I(DI_root_bodyadr); const int* rnum = I(DI_root_bodynum);
    const float* mass = F(DF_body_mass); const float* rinv = F(DF_root_invmass);
    int nroot = dim(DD_nroot);
    for (int r = tid; r < nroot; r += T) {
      V3 acc = v3(0, 0, 0);
      int b0 = radr[r], n = rnum[r];
      for (int b = b0; b < b0 + n; b++) acc = acc + ld3(w.xipos + 3 * b) * mass[b];
      if (rinv[r] > 0.f) acc = acc * rinv[r]; else acc = ld3(w.xipos + 3 * b0);
      st3(w.rootcom + 3 * r, acc);
    }
    sync();
    const int* ridx = I(DI_body_rootidx); const float* imat = F(DF_body_imat); const float* inertia = F(DF_body_inertia);
    int nb = dim(DD_nbody), nv = dim(DD_nv);
    for (int b = 1 + tid; b < nb; b += T) {
      V3 dif = ld3(w.xipos + 3 * b) - ld3(w.rootcom + 3 * ridx[b]);
      // ximat = xmat * imat
      const float* xm = w.xmat + 9 * b; const float* im = imat + 9 * b;
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
      float* ci = w.cinert + 10 * b;
      ci[0] = Ixx + m * (d2 - dif.x * dif.x); ci[1] = Iyy + m * (d2 - dif.y * dif.y); ci[2] = Izz + m * (d2 - dif.z * dif.z);
      ci[3] = Ixy - m * dif.x * dif.y; ci[4] = Ixz - m * dif.x * dif.z; ci[5] = Iyz - m * dif.y * dif.z;
      ci[6] = m * dif.x; ci[7] = m * dif.y; ci[8] = m * dif.z; ci[9] = m;
#pragma unroll
      for (int k = 0; k < 10; k++) w.crb[10 * b + k] = ci[k];
    }
    const int* isrot = I(DI_dof_isrot); const int* dbody = I(DI_dof_bodyid);
    for (int d = tid; d < nv; d += T) {
      if (isrot[d]) {
        V3 ax = ld3(w.cdof + 6 * d), com = ld3(w.rootcom + 3 * ridx[dbody[d]]);
        st3(w.cdof + 6 * d + 3, ld3(w.cdof + 6 * d + 3) + cross(ax, com));
      }
    }
    sync();
  }

  // ---- forward velocity / bias-acceleration pass (mj_comVel + first half of mj_rne), then local cfrc
  __device__ void vel_pass() {
    const int* parent = I(DI_body_parentid); const int* jadr = I(DI_body_jntadr); const int* jnum = I(DI_body_jntnum);
    const int* jtype = I(DI_jnt_type); const int* jd = I(DI_jnt_dofadr);
    const int* ladr = I(DI_level_adr); const int* lbody = I(DI_level_body);
    int maxdepth = dim(DD_maxdepth);
    for (int l = 1; l <= maxdepth; l++) {
      for (int idx = ladr[l] + tid; idx < ladr[l + 1]; idx += T) {
        int b = lbody[idx], p = parent[b], ja = jadr[b], jn = jnum[b];
        S6 cv = ld6(w.cvel + 6 * p), ca = ld6(w.cacc + 6 * p);
        for (int k = 0; k < jn; k++) {
          int j = ja + k, d = jd[j];
          if (jtype[j] == 0) {
#pragma unroll
            for (int a = 0; a < 3; a++) cv = cv + ld6(w.cdof + 6 * (d + a)) * w.qvel[d + a];
            S6 cv0 = cv;
#pragma unroll
            for (int a = 3; a < 6; a++) {
              S6 cd = ld6(w.cdof + 6 * (d + a)); float qv = w.qvel[d + a];
              ca = ca + cross_motion(cv0, cd) * qv; cv = cv + cd * qv;
            }
          } else {
            S6 cd = ld6(w.cdof + 6 * d); float qv = w.qvel[d];
            ca = ca + cross_motion(cv, cd) * qv; cv = cv + cd * qv;
          }
        }
        st6(w.cvel + 6 * b, cv); st6(w.cacc + 6 * b, ca);
        const float* ci = w.cinert + 10 * b;
        S6 f = mul_inert(ci, ca) + cross_force(cv, mul_inert(ci, cv));
        st6(w.cfrc + 6 * b, f);
      }
      sync();
    }
  }

  // ---- fused backward pass: composite inertias (mj_crb) and subtree bias forces (second half of mj_rne)
  __device__ void backward_pass() {
    const int* ladr = I(DI_level_adr); const int* lbody = I(DI_level_body);
    const int* cadr = I(DI_body_childadr); const int* cnum = I(DI_body_childnum); const int* child = I(DI_body_child);
    int maxdepth = dim(DD_maxdepth);
    for (int l = maxdepth - 1; l >= 1; l--) {
      for (int idx = ladr[l] + tid; idx < ladr[l + 1]; idx += T) {
        int b = lbody[idx], n = cnum[b];
        if (!n) continue;
        float acc[16];
#pragma unroll
        for (int k = 0; k < 10; k++) acc[k] = w.crb[10 * b + k];
#pragma unroll
        for (int k = 0; k < 6; k++) acc[10 + k] = w.cfrc[6 * b + k];
        for (int c = 0; c < n; c++) {
          int cb = child[cadr[b] + c];
#pragma unroll
          for (int k = 0; k < 10; k++) acc[k] += w.crb[10 * cb + k];
#pragma unroll
          for (int k = 0; k < 6; k++) acc[10 + k] += w.cfrc[6 * cb + k];
        }
#pragma unroll
        for (int k = 0; k < 10; k++) w.crb[10 * b + k] = acc[k];
#pragma unroll
        for (int k = 0; k < 6; k++) w.cfrc[6 * b + k] = acc[10 + k];
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
    for (int i = tid; i < nv; i += T) {
      S6 cd = ld6(w.cdof + 6 * i);
      S6 buf = mul_inert(w.crb + 10 * dbody[i], cd);
      int a = madr[i], n = ddepth[i];
      w.M[a] = dot6(cd, buf) + arm[i];
      for (int k = 1; k <= n; k++) w.M[a + k] = dot6(ld6(w.cdof + 6 * mcol[a + k]), buf);
      float qb = dot6(cd, ld6(w.cfrc + 6 * dbody[i]));
      float f = -damp[i] * w.qvel[i] - qb + w.qapp[i];
      int j = djnt[i];
      if (jtype[j] >= 2) { float k = stiff[j]; if (k != 0.f) f -= k * (w.qpos[jq[j]] - qspring[jq[j]]); }
      for (int u = 0; u < anum[i]; u++) {
        int ac = dofact[aadr[i] + u];
        float c = w.ctrl[ac];
        if (climited[ac]) c = clampf(c, crange[2 * ac], crange[2 * ac + 1]);
        float g = gear[ac];
        float af = gain[ac] * c + bias[3 * ac] + bias[3 * ac + 1] * g * w.qpos[jq[j]] + bias[3 * ac + 2] * g * w.qvel[i];
        if (flimited[ac]) af = clampf(af, frange[2 * ac], frange[2 * ac + 1]);
        f += g * af;
      }
      w.qfs[i] = f;
    }
    sync();
  }

  // ---- sparse L'DL factorisation of (M + hdamp*diag(damping)), one kinematic tree per warp (mj_factorM)
  __device__ void factor(float hdamp) {
    const int* madr = I(DI_dof_Madr); const int* ddepth = I(DI_dof_depth); const int* mcol = I(DI_Mcol);
    const int* tadr = I(DI_tree_dofadr); const int* tnum = I(DI_tree_dofnum); const float* damp = F(DF_dof_damping);
    int nM = dim(DD_nM), nv = dim(DD_nv), ntree = dim(DD_ntree);
    for (int k = tid; k < nM; k += T) w.LD[k] = w.M[k];
    sync();
    if (hdamp != 0.f) {
      for (int i = tid; i < nv; i += T) w.LD[madr[i]] += hdamp * damp[i];
      sync();
    }
    for (int t = warp; t < ntree; t += W) {
      int d0 = tadr[t], dn = tnum[t];
      for (int k = d0 + dn - 1; k >= d0; k--) {
        int ak = madr[k], dk = ddepth[k];
        float inv = 1.0f / w.LD[ak];
        // lane m (1..dk) owns row a_m = m-th ancestor of k
        for (int m = 1 + lane; m <= dk; m += 32) {
          int am = mcol[ak + m], aa = madr[am];
          float tk = w.LD[ak + m] * inv;
          for (int n = m; n <= dk; n++) w.LD[aa + (n - m)] -= tk * w.LD[ak + n];
        }
        __syncwarp();
        for (int m = 1 + lane; m <= dk; m += 32) w.LD[ak + m] *= inv;
        if (lane == 0) w.invD[k] = inv;
        __syncwarp();
      }
    }
    sync();
  }

  // ---- x <- M^-1 x with the factor above; level-synchronous over dof depth (mj_solveLD)
  __device__ void solve(float* x) {
    const int* madr = I(DI_dof_Madr); const int* ddepth = I(DI_dof_depth); const int* mcol = I(DI_Mcol);
    const int* dladr = I(DI_dlevel_adr); const int* dldof = I(DI_dlevel_dof);
    const int* dadr = I(DI_dof_descadr); const int* dnum = I(DI_dof_descnum); const int* ddof = I(DI_desc_dof);
    const int* dmadr = I(DI_desc_madr);
    int maxd = dim(DD_maxdofdepth), nv = dim(DD_nv);
    for (int l = maxd - 1; l >= 0; l--) {
      for (int idx = dladr[l] + tid; idx < dladr[l + 1]; idx += T) {
        int j = dldof[idx]; float s = x[j];
        for (int k = 0; k < dnum[j]; k++) s -= w.LD[dmadr[dadr[j] + k]] * x[ddof[dadr[j] + k]];
        x[j] = s;
      }
      sync();
    }
    // NOTE: the gather above needs every descendant final, which holds because deeper levels were finished first
    for (int i = tid; i < nv; i += T) x[i] *= w.invD[i];
    sync();
    for (int l = 1; l <= maxd; l++) {
      for (int idx = dladr[l] + tid; idx < dladr[l + 1]; idx += T) {
        int i = dldof[idx], a = madr[i], n = ddepth[i]; float s = x[i];
        for (int k = 1; k <= n; k++) s -= w.LD[a + k] * x[mcol[a + k]];
        x[i] = s;
      }
      sync();
    }
  }

  // ---- B.4 narrow phase, one candidate pair per thread, raw contacts into the arena, ordered compaction
  struct Raw { float dist; V3 pos, n, t; };

  __device__ __forceinline__ void geom_pose(int cg, V3& pos, float* mat) const {
    const int* cgbody = I(DI_cg_body); const float* cgpos = F(DF_cg_pos); const float* cgmat = F(DF_cg_mat);
    int b = cgbody[cg];
    const float* xm = w.xmat + 9 * b; const float* lm = cgmat + 9 * cg;
    pos = ld3(w.xpos + 3 * b) + mulmat(xm, ld3(cgpos + 3 * cg));
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) mat[3 * r + c] = xm[3 * r] * lm[c] + xm[3 * r + 1] * lm[3 + c] + xm[3 * r + 2] * lm[6 + c];
  }
  __device__ __forceinline__ int plane_sphere(Raw* out, V3 p1, V3 n, V3 p2, float radius, float margin) const {
    float cd = dot(p2 - p1, n);
    if (cd > margin + radius) return 0;
    out->dist = cd - radius; out->pos = p2 + n * (-out->dist * 0.5f - radius); out->n = n; out->t = v3(0, 0, 0);
    return 1;
  }
  __device__ int collide(int t1, int t2, V3 p1, const float* m1, const float* s1, V3 p2, const float* m2, const float* s2,
                         float margin, Raw* out) const {
    if (t1 == 0) {
      V3 n = matcol(m1, 2);
      if (t2 == 2) return plane_sphere(out, p1, n, p2, s2[0], margin);
      if (t2 == 3) {
        V3 ax = matcol(m2, 2); int cnt = 0;
        if (plane_sphere(out + cnt, p1, n, p2 + ax * s2[1], s2[0], margin)) { out[cnt].t = ax; cnt++; }
        if (plane_sphere(out + cnt, p1, n, p2 - ax * s2[1], s2[0], margin)) { out[cnt].t = ax; cnt++; }
        return cnt;
      }
      if (t2 == 6) {
        float dist = dot(p2 - p1, n); int cnt = 0;
        for (int i = 0; i < 8; i++) {
          V3 v = v3((i & 1) ? s2[0] : -s2[0], (i & 2) ? s2[1] : -s2[1], (i & 4) ? s2[2] : -s2[2]);
          V3 cn = mulmat(m2, v);
          float ld = dot(n, cn);
          if (dist + ld > margin || ld > 0.f) continue;
          out[cnt].dist = dist + ld; out[cnt].pos = cn + p2 - n * (out[cnt].dist * 0.5f); out[cnt].n = n; out[cnt].t = v3(0, 0, 0);
          if (++cnt >= 4) return 4;
        }
        return cnt;
      }
    }
    return 0;
  }
  __device__ void collision(unsigned long long* counters) {
    const int* pc1 = I(DI_pair_cg1); const int* pc2 = I(DI_pair_cg2); const int* pprm = I(DI_pair_prm);
    const int* praw = I(DI_pair_rawadr); const int* pmax = I(DI_pair_maxcon); const int* cgtype = I(DI_cg_type);
    const float* cgsize = F(DF_cg_size); const float* cgrb = F(DF_cg_rbound); const float* prm = F(DF_prm);
    int npair = dim(DD_npair);
    // raw slots: [count(int) per pair | 10 floats per raw contact] in the arena
    int* rcount = (int*)w.arena; float* rdata = w.arena + r4(npair);
    for (int p = tid; p < npair; p += T) {
      int g1 = pc1[p], g2 = pc2[p]; float margin = prm[B2DEV_PRM_STRIDE * pprm[p]];
      V3 p1, p2; float m1[9], m2[9];
      geom_pose(g1, p1, m1); geom_pose(g2, p2, m2);
      int n = 0; Raw raw[8];
      bool cull;
      if (cgtype[g1] == 0) cull = dot(p2 - p1, matcol(m1, 2)) > cgrb[g2] + margin;
      else { V3 d = p2 - p1; float bd = cgrb[g1] + cgrb[g2] + margin; cull = dot(d, d) > bd * bd; }
      if (!cull) n = collide(cgtype[g1], cgtype[g2], p1, m1, cgsize + 3 * g1, p2, m2, cgsize + 3 * g2, margin, raw);
      rcount[p] = n;
      float* dst = rdata + 10 * praw[p];
      for (int k = 0; k < n && k < pmax[p]; k++) {
        dst[10 * k] = raw[k].dist; st3(dst + 10 * k + 1, raw[k].pos); st3(dst + 10 * k + 4, raw[k].n); st3(dst + 10 * k + 7, raw[k].t);
      }
    }
    sync();
    // ordered compaction by warp 0 (pair order == MuJoCo contact order)
    if (warp == 0) {
      int base = 0, dropped = 0;
      for (int p0 = 0; p0 < npair; p0 += 32) {
        int p = p0 + lane; int n = (p < npair) ? rcount[p] : 0;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        int start = base + incl - n;
        for (int k = 0; k < n; k++) {
          int c = start + k;
          if (c >= conCap) { dropped++; continue; }
          const float* src = rdata + 10 * (praw[p] + k);
          float* dst = w.con + B2_CON_STRIDE * c;
          dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
          // frame: normal, then orthogonalised tangent hint (mju_makeFrame)
          V3 nrm = normalized(ld3(src + 4)); V3 t = ld3(src + 7);
          if (norm(t) < 0.5f) { t = (nrm.y < 0.5f && nrm.y > -0.5f) ? v3(0, 1, 0) : v3(0, 0, 1); }
          t = t - nrm * dot(nrm, t); t = normalized(t);
          V3 t2 = cross(nrm, t);
          st3(dst + 4, nrm); st3(dst + 7, t); st3(dst + 10, t2);
          dst[13] = __int_as_float(p); dst[14] = 0.f; dst[15] = 0.f;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
      }
      dropped = (int)warp_sum((float)dropped);
      if (lane == 0) {
        w.misc[MISC_NCON] = base < conCap ? base : conCap;
        if (dropped && counters) atomicAdd(&counters[CTR_CON_DROPPED], (unsigned long long)dropped);
      }
    }
    sync();
  }
  int conCap, rowCap, arenaFloats;

  // ---- B.5 rows: joint limits then pyramidal contacts, stably partitioned by island
  __device__ void make_rows(unsigned long long* counters) {
    const int* limj = I(DI_lim_jnt); const int* jq = I(DI_jnt_qposadr); const int* jd = I(DI_jnt_dofadr);
    const int* disl = I(DI_dof_island); const int* bisl = I(DI_body_island); const int* cgbody = I(DI_cg_body);
    const int* pc1 = I(DI_pair_cg1); const int* pc2 = I(DI_pair_cg2);
    const float* jrange = F(DF_jnt_range); const float* jmargin = F(DF_jnt_margin);
    int nlim = dim(DD_nlim), nisl = dim(DD_nisland), ncon = w.misc[MISC_NCON];
    if (tid < B2_MAX_ISLANDS) w.isl_n[tid] = 0;
    sync();
    const int maxrows = 32 * B2_PGS_S;
    if (warp == 0) {
      int dropped = 0;
      for (int c0 = 0; c0 < nlim; c0 += 32) {
        int k = c0 + lane; bool valid = k < nlim;
        int j = valid ? limj[k] : 0;
        float q = w.qpos[jq[j]], mg = jmargin[j];
        bool lo = valid && (q - jrange[2 * j] < mg), hi = valid && (jrange[2 * j + 1] - q < mg);
        int isl = valid ? disl[jd[j]] : -1 - lane;
        unsigned peers = __match_any_sync(0xffffffffu, isl);
        unsigned lomask = __ballot_sync(0xffffffffu, lo), himask = __ballot_sync(0xffffffffu, hi);
        unsigned lt = (1u << lane) - 1u;
        int before = __popc(lomask & peers & lt) + __popc(himask & peers & lt);
        int total = __popc(lomask & peers) + __popc(himask & peers);
        int base = valid ? w.isl_n[isl] : 0;
        __syncwarp();
        int r0 = base + before;
        if (valid) {
          int rl = -1, rh = -1;
          if (lo) { if (r0 < maxrows) rl = r0; else dropped++; r0++; }
          if (hi) { if (r0 < maxrows) rh = r0; else dropped++; }
          w.lim_row[2 * k] = rl; w.lim_row[2 * k + 1] = rh;
          if ((peers & lt) == 0 && total) w.isl_n[isl] = min(base + total, maxrows);
        }
        __syncwarp();
      }
      for (int c0 = 0; c0 < ncon; c0 += 32) {
        int c = c0 + lane; bool valid = c < ncon;
        int isl = -1 - lane;
        if (valid) {
          int p = __float_as_int(w.con[B2_CON_STRIDE * c + 13]);
          int i2 = bisl[cgbody[pc2[p]]], i1 = bisl[cgbody[pc1[p]]];
          isl = i2 >= 0 ? i2 : i1;
        }
        unsigned peers = __match_any_sync(0xffffffffu, isl);
        unsigned lt = (1u << lane) - 1u;
        int before = 4 * __popc(peers & lt), total = 4 * __popc(peers);
        int base = valid ? w.isl_n[isl] : 0;
        __syncwarp();
        if (valid) {
          int r0 = base + before;
          if (r0 + 4 <= maxrows) w.con_row[c] = r0; else { w.con_row[c] = -1; dropped++; }
          if ((peers & lt) == 0) w.isl_n[isl] = base + min(total, base < maxrows ? ((maxrows - base) / 4) * 4 : 0);
        }
        __syncwarp();
      }
      // island bases and arena carve-up
      if (lane == 0) {
        int adr = 0, used = 0, ovf = 0;
        const int* inum = I(DI_island_dofnum);
        for (int k = 0; k < nisl; k++) {
          int n = w.isl_n[k]; n = min(n, maxrows);
          // drop a partially filled pyramid at the cap
          int ldj = inum[k] | 1;
          int need = n * ldj + n * n;
          if (adr + n > rowCap || used + need > arenaFloats - scratchFloats()) { n = 0; need = 0; ovf++; }
          w.isl_n[k] = n; w.isl_adr[k] = adr; w.isl_ldj[k] = ldj; w.isl_J[k] = used; w.isl_A[k] = used + n * ldj;
          adr += n; used += need;
        }
        w.isl_adr[nisl] = adr; w.misc[MISC_NEFC] = adr; w.misc[MISC_ARENA_USED] = used;
        if (counters) {
          if (ovf) atomicAdd(&counters[CTR_ARENA_OVERFLOW], (unsigned long long)ovf);
        }
      }
      dropped = (int)warp_sum((float)dropped);
      if (lane == 0 && dropped && counters) atomicAdd(&counters[CTR_ROW_DROPPED], (unsigned long long)dropped);
    }
    sync();
  }
  __device__ __forceinline__ int scratchFloats() const {
    // per-warp column scratch for the A build: 33 * max island dof span
    int m = 0; const int* inum = I(DI_island_dofnum);
    for (int k = 0; k < dim(DD_nisland); k++) m = max(m, inum[k]);
    return W * 33 * m;
  }

  // ---- fill J (island-dense), per-row parameters, aref, b, warm-start force
  __device__ void fill_rows() {
    const int* limj = I(DI_lim_jnt); const int* jd = I(DI_jnt_dofadr); const int* jq = I(DI_jnt_qposadr);
    const int* disl = I(DI_dof_island); const int* bisl = I(DI_body_island); const int* cgbody = I(DI_cg_body);
    const int* pc1 = I(DI_pair_cg1); const int* pc2 = I(DI_pair_cg2); const int* pprm = I(DI_pair_prm);
    const int* iadr = I(DI_island_dofadr); const int* inum = I(DI_island_dofnum); const int* dbody = I(DI_dof_bodyid);
    const int* ridx = I(DI_body_rootidx); const int* cmask = I(DI_body_chainmask);
    const float* jrange = F(DF_jnt_range); const float* jmargin = F(DF_jnt_margin); const float* jsol = F(DF_jnt_solprm);
    const float* dinvw = F(DF_dof_invweight0); const float* binvw = F(DF_body_invweight0); const float* prm = F(DF_prm);
    int nlim = dim(DD_nlim), nisl = dim(DD_nisland), ncon = w.misc[MISC_NCON], nmw = dim(DD_nmaskw);
    float timestep = P.opt[DO_timestep], impratio = P.opt[DO_impratio];
    // row_info: limits  -> (joint << 2) | side ; contacts -> 0x40000000 | (contact << 2) | dir
    for (int k = tid; k < 2 * nlim; k += T) {
      int r = w.lim_row[k]; if (r < 0) continue;
      int j = limj[k >> 1], isl = disl[jd[j]];
      if (r >= w.isl_n[isl]) continue;
      w.row_info[w.isl_adr[isl] + r] = (j << 2) | (k & 1);
    }
    for (int c = tid; c < ncon; c += T) {
      int r = w.con_row[c]; if (r < 0) continue;
      int p = __float_as_int(w.con[B2_CON_STRIDE * c + 13]);
      int i2 = bisl[cgbody[pc2[p]]], i1 = bisl[cgbody[pc1[p]]]; int isl = i2 >= 0 ? i2 : i1;
      if (r + 4 > w.isl_n[isl]) { w.con_row[c] = -1; continue; }
      for (int d = 0; d < 4; d++) w.row_info[w.isl_adr[isl] + r + d] = 0x40000000 | (c << 2) | d;
    }
    sync();
    // J entries
    for (int k = 0; k < nisl; k++) {
      int n = w.isl_n[k]; if (!n) continue;
      int d0 = iadr[k], nd = inum[k], ldj = w.isl_ldj[k]; float* J = w.arena + w.isl_J[k];
      int e0 = w.isl_adr[k];
      for (int item = tid; item < n * nd; item += T) {
        int i = item / nd, c = item - i * nd, d = d0 + c;
        int info = w.row_info[e0 + i]; float val = 0.f;
        if (info & 0x40000000) {
          int ci = (info >> 2) & 0x0fffffff, dir = info & 3;
          const float* con = w.con + B2_CON_STRIDE * ci;
          int p = __float_as_int(con[13]);
          int b1 = cgbody[pc1[p]], b2 = cgbody[pc2[p]];
          int in1 = (cmask[b1 * nmw + (d >> 5)] >> (d & 31)) & 1, in2 = (cmask[b2 * nmw + (d >> 5)] >> (d & 31)) & 1;
          int sgn = in2 - in1;
          if (sgn) {
            const float* pr = prm + B2DEV_PRM_STRIDE * pprm[p];
            float mu = pr[2 + (dir >> 1)];
            V3 nrm = ld3(con + 4), tv = ld3(con + 7 + 3 * (dir >> 1));
            V3 dv = nrm + tv * ((dir & 1) ? -mu : mu);
            S6 cd = ld6(w.cdof + 6 * d);
            V3 off = ld3(con + 1) - ld3(w.rootcom + 3 * ridx[dbody[d]]);
            val = (float)sgn * dot(dv, cd.l + cross(cd.a, off));
          }
        } else {
          int j = info >> 2;
          if (jd[j] == d) val = (info & 1) ? -1.f : 1.f;
        }
        J[i * ldj + c] = val;
      }
    }
    sync();
    // per-row parameters
    int nefc = w.misc[MISC_NEFC];
    for (int e = tid; e < nefc; e += T) {
      int info = w.row_info[e];
      float pos, margin, da, solref0, solref1; const float* simp; int isl; float mu0 = 0.f; bool iscon = info & 0x40000000;
      if (iscon) {
        int ci = (info >> 2) & 0x0fffffff;
        const float* con = w.con + B2_CON_STRIDE * ci; int p = __float_as_int(con[13]);
        const float* pr = prm + B2DEV_PRM_STRIDE * pprm[p];
        int b1 = cgbody[pc1[p]], b2 = cgbody[pc2[p]];
        isl = bisl[b2] >= 0 ? bisl[b2] : bisl[b1];
        pos = con[0]; margin = pr[0] - pr[1]; mu0 = pr[2];
        float tran = binvw[2 * b1] + binvw[2 * b2];
        da = tran + mu0 * mu0 * tran;   // first row of the pyramid sets R for all of its rows
        solref0 = pr[7]; solref1 = pr[8]; simp = pr + 9;
      } else {
        int j = info >> 2, side = info & 1; float q = w.qpos[jq[j]];
        isl = disl[jd[j]];
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
          else if (x <= mid) y = powf(x, power) / powf(mid, power - 1.f);
          else y = 1.f - powf(1.f - x, power) / powf(1.f - mid, power - 1.f);
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
      // J row products with qvel, qacc_smooth, qacc_warmstart
      int d0 = iadr[isl], nd = inum[isl], ldj = w.isl_ldj[isl];
      const float* Jr = w.arena + w.isl_J[isl] + (e - w.isl_adr[isl]) * ldj;
      float vel = 0.f, ja = 0.f, jw = 0.f;
      for (int c = 0; c < nd; c++) { float jv = Jr[c]; vel = fmaf(jv, w.qvel[d0 + c], vel); ja = fmaf(jv, w.qas[d0 + c], ja); jw = fmaf(jv, w.warm[d0 + c], jw); }
      float aref = -B * vel - K * imp * (pos - margin);
      float D = 1.f / R;
      w.row_pos[e] = pos; w.row_margin[e] = margin; w.row_R[e] = R; w.row_D[e] = D; w.row_aref[e] = aref;
      w.row_b[e] = ja - aref;
      float jar = jw - aref;
      w.row_f[e] = jar < 0.f ? -D * jar : 0.f;
    }
    sync();
  }

  // ---- A = J M^-1 J' + R per island (mj_projectConstraint); one island per warp, 32 columns at a time
  __device__ void build_A() {
    const int* madr = I(DI_dof_Madr); const int* ddepth = I(DI_dof_depth); const int* mcol = I(DI_Mcol);
    const int* iadr = I(DI_island_dofadr); const int* inum = I(DI_island_dofnum);
    int nisl = dim(DD_nisland);
    int maxspan = 0; for (int k = 0; k < nisl; k++) maxspan = max(maxspan, inum[k]);
    float* scratch = w.arena + arenaFloats - (warp + 1) * 33 * maxspan;
    for (int k = warp; k < nisl; k += W) {
      int n = w.isl_n[k]; if (!n) continue;
      int d0 = iadr[k], nd = inum[k], ldj = w.isl_ldj[k], e0 = w.isl_adr[k];
      const float* J = w.arena + w.isl_J[k]; float* A = w.arena + w.isl_A[k];
      for (int j0 = 0; j0 < n; j0 += 32) {
        int j = j0 + lane; bool valid = j < n;
        float* x = scratch + lane;   // column, stride 33
        for (int c = 0; c < nd; c++) x[33 * c] = valid ? J[j * ldj + c] : 0.f;
        // x <- L^-T x
        for (int i = nd - 1; i >= 0; i--) {
          float xi = x[33 * i]; int a = madr[d0 + i], dn = ddepth[d0 + i];
          for (int m = 1; m <= dn; m++) { int cc = mcol[a + m] - d0; x[33 * cc] -= w.LD[a + m] * xi; }
        }
        for (int i = 0; i < nd; i++) x[33 * i] *= w.invD[d0 + i];
        for (int i = 0; i < nd; i++) {
          float s = x[33 * i]; int a = madr[d0 + i], dn = ddepth[d0 + i];
          for (int m = 1; m <= dn; m++) { int cc = mcol[a + m] - d0; s -= w.LD[a + m] * x[33 * cc]; }
          x[33 * i] = s;
        }
        // A[i][j] = J_i . x
        for (int i = 0; i < n; i++) {
          const float* Ji = J + i * ldj; float s = 0.f;
          for (int c = 0; c < nd; c++) s = fmaf(Ji[c], x[33 * c], s);
          if (valid) A[i * n + j] = s + (i == j ? w.row_R[e0 + i] : 0.f);
        }
        __syncwarp();
      }
    }
    sync();
  }

  // ---- PGS (mj_solPGS restated; row order = MuJoCo's within each island, islands are exactly decoupled)
  __device__ void solve_pgs(unsigned long long* counters) {
    int nisl = dim(DD_nisland), iters = dim(DD_iterations);
    float scale = P.opt[DO_pgs_scale], tol = P.opt[DO_tolerance];
    // warm-start acceptance: cost(f) = 1/2 f'A f + f'b > 0 -> cold start (engine_forward.c warmstart())
    float cost = 0.f;
    for (int k = warp; k < nisl; k += W) {
      int n = w.isl_n[k]; if (!n) continue;
      int e0 = w.isl_adr[k]; const float* A = w.arena + w.isl_A[k];
      for (int i = lane; i < n; i += 32) {
        float s = 0.f;
        for (int j = 0; j < n; j++) s = fmaf(A[j * n + i], w.row_f[e0 + j], s);
        float fi = w.row_f[e0 + i];
        cost += fi * (0.5f * s + w.row_b[e0 + i]);
      }
    }
    cost = warp_sum(cost);
    if (lane == 0) w.red[warp] = cost;
    sync();
    float total = 0.f;
    for (int k = 0; k < W; k++) total += w.red[k];
    bool cold = total > 0.f;
    sync();
    int it = 0;
    // per-island register state is re-loaded each iteration (a warp may own several islands)
    for (int k = warp; k < nisl; k += W) {
      int n = w.isl_n[k]; if (!n) continue;
      int e0 = w.isl_adr[k];
      if (cold) for (int i = lane; i < n; i += 32) w.row_f[e0 + i] = 0.f;
    }
    __syncwarp();
    // residual r = A f + b kept in row_aref's slot? no: aref is still needed by nobody after b is formed -> reuse row_aref
    float* res = w.row_aref;
    for (int k = warp; k < nisl; k += W) {
      int n = w.isl_n[k]; if (!n) continue;
      int e0 = w.isl_adr[k]; const float* A = w.arena + w.isl_A[k];
      for (int i = lane; i < n; i += 32) {
        float s = w.row_b[e0 + i];
        if (!cold) for (int j = 0; j < n; j++) s = fmaf(A[j * n + i], w.row_f[e0 + j], s);
        res[e0 + i] = s;
      }
    }
    __syncwarp();
    for (; it < iters; it++) {
      float improvement = 0.f;
      for (int k = warp; k < nisl; k += W) {
        int n = w.isl_n[k]; if (!n) continue;
        int e0 = w.isl_adr[k]; const float* A = w.arena + w.isl_A[k];
        float f[B2_PGS_S], r[B2_PGS_S], ad[B2_PGS_S], ainv[B2_PGS_S];
#pragma unroll
        for (int s = 0; s < B2_PGS_S; s++) {
          int i = lane + 32 * s; bool v = i < n;
          f[s] = v ? w.row_f[e0 + i] : 0.f; r[s] = v ? res[e0 + i] : 0.f;
          ad[s] = v ? A[i * n + i] : 1.f; ainv[s] = 1.f / ad[s];
        }
#pragma unroll
        for (int s = 0; s < B2_PGS_S; s++) {
          int nn = min(32, n - 32 * s);
          for (int ii = 0; ii < nn; ii++) {
            float nf = fmaxf(0.f, f[s] - r[s] * ainv[s]);
            float dl = nf - f[s];
            float ch = dl * (0.5f * dl * ad[s] + r[s]);
            if (ch > 1e-10f) { dl = 0.f; ch = 0.f; }
            if (lane == ii) { improvement -= ch; f[s] += dl; }
            dl = __shfl_sync(0xffffffffu, dl, ii);
            if (dl != 0.f) {
              const float* Ar = A + (32 * s + ii) * n;
#pragma unroll
              for (int s2 = 0; s2 < B2_PGS_S; s2++) { int c = lane + 32 * s2; if (c < n) r[s2] = fmaf(Ar[c], dl, r[s2]); }
            }
          }
        }
#pragma unroll
        for (int s = 0; s < B2_PGS_S; s++) { int i = lane + 32 * s; if (i < n) { w.row_f[e0 + i] = f[s]; res[e0 + i] = r[s]; } }
      }
      improvement = warp_sum(improvement);
      if (lane == 0) w.red[8 + (it & 1) * 8 + warp] = improvement;
      sync();
      float tot = 0.f;
      for (int k = 0; k < W; k++) tot += w.red[8 + (it & 1) * 8 + k];
      if (tot * scale < tol) { it++; break; }
    }
    if (tid == 0) { w.misc[MISC_ITERS] = it; if (counters) atomicAdd(&counters[CTR_SOLVER_ITERS], (unsigned long long)it); }
    sync();
  }

  // ---- qfrc_constraint = J' f ; qacc = qacc_smooth + M^-1 qfrc_constraint
  __device__ void finish_constraint() {
    const int* disl = I(DI_dof_island); const int* iadr = I(DI_island_dofadr);
    int nv = dim(DD_nv);
    for (int d = tid; d < nv; d += T) {
      int k = disl[d]; int n = w.isl_n[k]; float s = 0.f;
      if (n) {
        int ldj = w.isl_ldj[k], e0 = w.isl_adr[k], c = d - iadr[k]; const float* J = w.arena + w.isl_J[k];
        for (int i = 0; i < n; i++) s = fmaf(J[i * ldj + c], w.row_f[e0 + i], s);
      }
      w.qfc[d] = s; w.qacc[d] = s;
    }
    sync();
    solve(w.qacc);
    for (int d = tid; d < nv; d += T) w.qacc[d] += w.qas[d];
    sync();
  }

  // ---- mj_forward
  __device__ void forward(unsigned long long* counters) {
    int nv = dim(DD_nv);
    kinematics(); com_pos(); vel_pass(); backward_pass(); mass_and_smooth();
    factor(0.f);
    for (int d = tid; d < nv; d += T) w.qas[d] = w.qfs[d];
    sync();
    solve(w.qas);
    collision(counters);
    make_rows(counters);
    if (w.misc[MISC_NEFC] > 0) {
      fill_rows(); build_A(); solve_pgs(counters); finish_constraint();
    } else {
      for (int d = tid; d < nv; d += T) { w.qacc[d] = w.qas[d]; w.qfc[d] = 0.f; }
      if (tid == 0) w.misc[MISC_ITERS] = 0;
      sync();
    }
  }

  __device__ void reset_data() {
    const float* q0 = F(DF_qpos0);
    int nq = dim(DD_nq), nv = dim(DD_nv), nu = dim(DD_nu);
    for (int i = tid; i < nq; i += T) w.qpos[i] = q0[i];
    for (int i = tid; i < nv; i += T) { w.qvel[i] = 0.f; w.warm[i] = 0.f; w.qapp[i] = 0.f; }
    for (int i = tid; i < nu; i += T) w.ctrl[i] = 0.f;
    sync();
  }
  __device__ bool bad_state(const float* x, int n) {
    int bad = 0;
    for (int i = tid; i < n; i += T) { float v = x[i]; if (!(v == v) || fabsf(v) > B2_MAXVAL) bad = 1; }
    if (T == 32) return __any_sync(0xffffffffu, bad);
    return __syncthreads_or(bad);
  }

  // ---- mj_step with the Euler integrator (implicit joint damping) -- SURVEY B.0 / B.7
  __device__ void step_euler(unsigned long long* counters) {
    float* time = w.time;
    const int* jtype = I(DI_jnt_type); const int* jq = I(DI_jnt_qposadr); const int* jd = I(DI_jnt_dofadr);
    int nq = dim(DD_nq), nv = dim(DD_nv), njnt = dim(DD_njnt); float h = P.opt[DO_timestep];
    if (bad_state(w.qpos, nq) | bad_state(w.qvel, nv)) {
      reset_data(); if (tid == 0) { *time = 0.f; if (counters) atomicAdd(&counters[CTR_NAN_RESET], 1ull); }
    }
    forward(counters);
    if (bad_state(w.qacc, nv)) {
      reset_data(); if (tid == 0) { *time = 0.f; if (counters) atomicAdd(&counters[CTR_NAN_RESET], 1ull); }
      forward(counters);
    }
    // (M + h*diag(damping)) qacc' = qfrc_smooth + qfrc_constraint
    factor(h);
    for (int d = tid; d < nv; d += T) w.tmp[d] = w.qfs[d] + w.qfc[d];
    sync();
    solve(w.tmp);
    for (int d = tid; d < nv; d += T) { w.qvel[d] += h * w.tmp[d]; w.warm[d] = w.qacc[d]; }
    sync();
    for (int j = tid; j < njnt; j += T) {
      int qa = jq[j], da = jd[j];
      if (jtype[j] == 0) {
        for (int k = 0; k < 3; k++) w.qpos[qa + k] += h * w.qvel[da + k];
        V3 om = ld3(w.qvel + da + 3); float n = norm(om);
        if (n >= B2_MINVAL) {
          Q4 q = qnormalize(qmul(ldq(w.qpos + qa + 3), axisangle(om * (1.f / n), h * n)));
          stq(w.qpos + qa + 3, q);
        }
      } else w.qpos[qa] += h * w.qvel[da];
    }
    if (tid == 0) { *time += h; if (counters) atomicAdd(&counters[CTR_SUBSTEPS], 1ull); }
    sync();
  }
};

}  // namespace b2
