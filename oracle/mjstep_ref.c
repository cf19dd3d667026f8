/* oracle/mjstep_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * fp64 CPU restatement of the `mujoco.mj_step` pipeline restricted to the feature subset the
 * seven reference tasks exercise.  It is the checker the CUDA path is compared
 * against; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product path never links, imports or falls back to this file.
 *
 * PARITY UNPINNED: the arithmetic being restated lives in the third-party `mujoco` wheel
 * (un-vendored; `mujoco>=2.3.0` / `>=3.0.0`, no lockfile: robotic_arm_assembly_env/setup.py:26,
 * humanoid_martial_arts_env/setup.py:28).  It can be imported neither in the authoring container
 * nor on the GPU box and the reference repository holds no golden vector for this path (0 asserts
 * in its tests), so this file restates MuJoCo's published algorithm (engine_forward.c,
 * engine_core_smooth.c, engine_collision_primitive.c, engine_core_constraint.c, engine_solver.c as of
 * MuJoCo 3.x) and is anchored on the reference's call sites:
 *     mj_step      quadruped_parkour_env/parkour_env.py:348,368  humanoid_dancing_env/dancing_env.py:812,849
 *                  humanoid_soccer_env/soccer_env.py:379,414     bipedal_rescue_env/rescue_env.py:398,432
 *                  humanoid_martial_arts_env/martial_arts_env.py:498  humanoid_construction_env/construction_env.py:595
 *                  robotic_arm_assembly_env/assembly_env.py:187,229
 *     mj_forward   humanoid_martial_arts_env/martial_arts_env.py:481
 *     mj_resetData parkour_env.py:321 and the six sibling reset() functions
 * What pins it instead: analytic known-answer tests (tests/test_oracle_kat.py) and an independent
 * dense-Jacobian mass matrix in the model compiler.
 *
 * Deliberate simplifications that do not change results beyond fp64 round-off: the joint-space
 * inertia is held dense and factorised by Cholesky (MuJoCo: sparse L'DL); efc_J and A_R are dense.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/b2_model_layout.h"

typedef double real;
#define MINVAL 1e-15
#define MAXVAL 1e10
#define MINIMP 0.0001
#define MAXIMP 0.9999

enum { CNSTR_LIMIT = 3, CNSTR_CONTACT_PYRAMIDAL = 5, CNSTR_CONTACT_FRICTIONLESS = 4 };

typedef struct {
  int d[32];
  real timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  int *ints; real *flts;
  const int *I[B2_N_INT_FIELDS];
  const real *F[B2_N_FLT_FIELDS];
} RefModel;

#define NQ (m->d[B2D_nq])
#define NV (m->d[B2D_nv])
#define NU (m->d[B2D_nu])
#define NBODY (m->d[B2D_nbody])
#define NJNT (m->d[B2D_njnt])
#define NGEOM (m->d[B2D_ngeom])
#define NSITE (m->d[B2D_nsite])
#define NPAIR (m->d[B2D_npair])
#define MI(f) (m->I[B2I_##f])
#define MF(f) (m->F[B2F_##f])

typedef struct {
  real dist, pos[3], frame[9], includemargin, friction[5], solref[2], solimp[5];
  int dim, geom1, geom2, efc_address, pair;
} RefContact;

typedef struct {
  real time;
  real *qpos, *qvel, *ctrl, *qfrc_applied, *xfrc_applied, *qacc, *qacc_warmstart;
  real *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *geom_xpos, *geom_xmat, *site_xpos, *site_xmat;
  real *subtree_com, *cinert, *crb, *cdof, *cdof_dot, *cvel, *cacc, *cfrc;
  real *M, *L, *qfrc_bias, *qfrc_passive, *qfrc_actuator, *qfrc_smooth, *qacc_smooth, *qfrc_constraint;
  real *actuator_force;
  int ncon, nefc, maxcon, maxefc, arcap;
  RefContact *contact;
  real *efc_J, *efc_pos, *efc_margin, *efc_D, *efc_R, *efc_aref, *efc_vel, *efc_b, *efc_force, *efc_AR;
  real *efc_diagApprox, *efc_KBIP;
  int *efc_type, *efc_id;
  int solver_iter, nwarn_bad, ncon_dropped;
  real *scratch; /* >= 8*nv + nv*nv */
  int disable_eulerdamp, disable_warmstart;
  int warmstart_once_per_step;   /* 0 (default): qacc_warmstart is saved at the end of every mj_fwdConstraint, as MuJoCo 3.x does; 1: once per mj_step */
} RefData;

/* ------------------------------------------------------------------ small vector helpers */
static inline real dot3(const real *a, const real *b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static inline void cross3(real *r, const real *a, const real *b) {
  real x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline real norm3(const real *a) { return sqrt(dot3(a, a)); }
static inline real normalize3(real *a) {
  real n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; } else { a[0] /= n; a[1] /= n; a[2] /= n; }
  return n;
}
static inline void mulmatvec3(real *r, const real *mat, const real *v) {
  real x = mat[0]*v[0] + mat[1]*v[1] + mat[2]*v[2];
  real y = mat[3]*v[0] + mat[4]*v[1] + mat[5]*v[2];
  real z = mat[6]*v[0] + mat[7]*v[1] + mat[8]*v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulquat(real *r, const real *a, const real *b) {
  real w = a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3];
  real x = a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2];
  real y = a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1];
  real z = a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void normalize4(real *q) {
  real n = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n; }
}
static inline void quat2mat(real *m, const real *q) {
  real q00 = q[0]*q[0], q11 = q[1]*q[1], q22 = q[2]*q[2], q33 = q[3]*q[3];
  real q01 = q[0]*q[1], q02 = q[0]*q[2], q03 = q[0]*q[3], q12 = q[1]*q[2], q13 = q[1]*q[3], q23 = q[2]*q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = 2*(q12 - q03); m[2] = 2*(q13 + q02);
  m[3] = 2*(q12 + q03); m[4] = q00 - q11 + q22 - q33; m[5] = 2*(q23 - q01);
  m[6] = 2*(q13 - q02); m[7] = 2*(q23 + q01); m[8] = q00 - q11 - q22 + q33;
}
static inline void rotvecquat(real *r, const real *v, const real *q) {
  real mat[9]; quat2mat(mat, q); mulmatvec3(r, mat, v);
}
static inline void axisangle2quat(real *q, const real *axis, real angle) {
  real s = sin(angle * 0.5);
  q[0] = cos(angle * 0.5); q[1] = axis[0]*s; q[2] = axis[1]*s; q[3] = axis[2]*s;
}
/* quat <- quat * exp(h*omega/2), omega expressed in the local frame (mju_quatIntegrate) */
static void quat_integrate(real *quat, const real *vel, real h) {
  real ax[3] = {vel[0], vel[1], vel[2]};
  real n = norm3(ax);
  if (n < MINVAL) return;
  ax[0] /= n; ax[1] /= n; ax[2] /= n;
  real qr[4], out[4];
  axisangle2quat(qr, ax, h * n);
  mulquat(out, quat, qr);
  normalize4(out);
  memcpy(quat, out, sizeof(out));
}

/* ------------------------------------------------------------------ model / data lifetime */
RefModel *ref_model_create(const int *ints, int n_ints, const double *flts, int n_flts) {
  if (ints[0] != B2_MAGIC || ints[1] != B2_N_INT_FIELDS || ints[2] != B2_N_FLT_FIELDS) return NULL;
  RefModel *m = (RefModel *)calloc(1, sizeof(RefModel));
  m->ints = (int *)malloc(sizeof(int) * n_ints); memcpy(m->ints, ints, sizeof(int) * n_ints);
  m->flts = (real *)malloc(sizeof(real) * n_flts); memcpy(m->flts, flts, sizeof(real) * n_flts);
  for (int k = 0; k < B2_N_INT_FIELDS; k++) m->I[k] = m->ints + B2_INT_OFF(m->ints, k);
  for (int k = 0; k < B2_N_FLT_FIELDS; k++) m->F[k] = m->flts + B2_FLT_OFF(m->ints, k);
  int nd = B2_INT_CNT(m->ints, B2I_dims);
  for (int k = 0; k < nd && k < 32; k++) m->d[k] = m->I[B2I_dims][k];
  const real *o = m->F[B2F_opt];
  m->timestep = o[B2O_timestep]; m->gravity[0] = o[B2O_gravity_x]; m->gravity[1] = o[B2O_gravity_y];
  m->gravity[2] = o[B2O_gravity_z]; m->tolerance = o[B2O_tolerance]; m->ls_tolerance = o[B2O_ls_tolerance];
  m->impratio = o[B2O_impratio]; m->meaninertia = o[B2O_meaninertia];
  return m;
}
void ref_model_destroy(RefModel *m) { if (m) { free(m->ints); free(m->flts); free(m); } }

static real *zalloc(size_t n) { return (real *)calloc(n ? n : 1, sizeof(real)); }

void ref_reset_data(const RefModel *m, RefData *d);

RefData *ref_data_create(const RefModel *m) {
  RefData *d = (RefData *)calloc(1, sizeof(RefData));
  int nq = NQ, nv = NV, nb = NBODY;
  d->qpos = zalloc(nq); d->qvel = zalloc(nv); d->ctrl = zalloc(NU); d->qfrc_applied = zalloc(nv);
  d->xfrc_applied = zalloc(6 * nb); d->qacc = zalloc(nv); d->qacc_warmstart = zalloc(nv);
  d->xpos = zalloc(3 * nb); d->xquat = zalloc(4 * nb); d->xmat = zalloc(9 * nb); d->xipos = zalloc(3 * nb);
  d->ximat = zalloc(9 * nb); d->xanchor = zalloc(3 * NJNT); d->xaxis = zalloc(3 * NJNT);
  d->geom_xpos = zalloc(3 * NGEOM); d->geom_xmat = zalloc(9 * NGEOM);
  d->site_xpos = zalloc(3 * NSITE); d->site_xmat = zalloc(9 * NSITE);
  d->subtree_com = zalloc(3 * nb); d->cinert = zalloc(10 * nb); d->crb = zalloc(10 * nb);
  d->cdof = zalloc(6 * nv); d->cdof_dot = zalloc(6 * nv); d->cvel = zalloc(6 * nb); d->cacc = zalloc(6 * nb);
  d->cfrc = zalloc(6 * nb);
  d->M = zalloc(nv * nv); d->L = zalloc(nv * nv); d->qfrc_bias = zalloc(nv); d->qfrc_passive = zalloc(nv);
  d->qfrc_actuator = zalloc(nv); d->qfrc_smooth = zalloc(nv); d->qacc_smooth = zalloc(nv);
  d->qfrc_constraint = zalloc(nv); d->actuator_force = zalloc(NU);
  d->maxcon = m->d[B2D_maxcon] < 2000 ? m->d[B2D_maxcon] : 2000;
  d->contact = (RefContact *)calloc(d->maxcon ? d->maxcon : 1, sizeof(RefContact));
  d->maxefc = 2 * m->d[B2D_nlimited] + 10 * d->maxcon;
  int ne = d->maxefc ? d->maxefc : 1;
  d->efc_pos = zalloc(ne); d->efc_margin = zalloc(ne); d->efc_D = zalloc(ne); d->efc_R = zalloc(ne);
  d->efc_aref = zalloc(ne); d->efc_vel = zalloc(ne); d->efc_b = zalloc(ne); d->efc_force = zalloc(ne);
  d->efc_diagApprox = zalloc(ne); d->efc_KBIP = zalloc(4 * ne);
  d->efc_type = (int *)calloc(ne, sizeof(int)); d->efc_id = (int *)calloc(ne, sizeof(int));
  d->arcap = 0; d->efc_J = NULL; d->efc_AR = NULL;
  d->scratch = zalloc(16 * nv + 2 * nv * nv + 64);
  ref_reset_data(m, d);
  return d;
}
void ref_data_destroy(RefData *d) {
  if (!d) return;
  real **p[] = {&d->qpos, &d->qvel, &d->ctrl, &d->qfrc_applied, &d->xfrc_applied, &d->qacc, &d->qacc_warmstart,
                &d->xpos, &d->xquat, &d->xmat, &d->xipos, &d->ximat, &d->xanchor, &d->xaxis, &d->geom_xpos,
                &d->geom_xmat, &d->site_xpos, &d->site_xmat, &d->subtree_com, &d->cinert, &d->crb, &d->cdof,
                &d->cdof_dot, &d->cvel, &d->cacc, &d->cfrc, &d->M, &d->L, &d->qfrc_bias, &d->qfrc_passive,
                &d->qfrc_actuator, &d->qfrc_smooth, &d->qacc_smooth, &d->qfrc_constraint, &d->actuator_force,
                &d->efc_J, &d->efc_pos, &d->efc_margin, &d->efc_D, &d->efc_R, &d->efc_aref, &d->efc_vel,
                &d->efc_b, &d->efc_force, &d->efc_AR, &d->efc_diagApprox, &d->efc_KBIP, &d->scratch};
  for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); i++) free(*p[i]);
  free(d->contact); free(d->efc_type); free(d->efc_id); free(d);
}

/* mj_resetData */
void ref_reset_data(const RefModel *m, RefData *d) {
  memcpy(d->qpos, MF(qpos0), sizeof(real) * NQ);
  memset(d->qvel, 0, sizeof(real) * NV); memset(d->ctrl, 0, sizeof(real) * NU);
  memset(d->qfrc_applied, 0, sizeof(real) * NV); memset(d->xfrc_applied, 0, sizeof(real) * 6 * NBODY);
  memset(d->qacc, 0, sizeof(real) * NV); memset(d->qacc_warmstart, 0, sizeof(real) * NV);
  d->time = 0; d->ncon = 0; d->nefc = 0; d->solver_iter = 0;
}

/* ------------------------------------------------------------------ B.1 position stage */
static void kinematics(const RefModel *m, RefData *d) {
  const int *parent = MI(body_parentid), *jntadr = MI(body_jntadr), *jntnum = MI(body_jntnum);
  const int *jtype = MI(jnt_type), *qadr = MI(jnt_qposadr);
  d->xpos[0] = d->xpos[1] = d->xpos[2] = 0;
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  for (int j = 0; j < NJNT; j++) if (jtype[j] == B2_JNT_FREE) normalize4(d->qpos + qadr[j] + 3);
  for (int b = 1; b < NBODY; b++) {
    real *xp = d->xpos + 3*b, *xq = d->xquat + 4*b;
    int ja = jntadr[b], jn = jntnum[b], p = parent[b];
    if (jn == 1 && jtype[ja] == B2_JNT_FREE) {
      const real *q = d->qpos + qadr[ja];
      xp[0] = q[0]; xp[1] = q[1]; xp[2] = q[2];
      xq[0] = q[3]; xq[1] = q[4]; xq[2] = q[5]; xq[3] = q[6];
      memcpy(d->xanchor + 3*ja, xp, 3 * sizeof(real));
      rotvecquat(d->xaxis + 3*ja, MF(jnt_axis) + 3*ja, xq);
    } else {
      real t[3];
      mulmatvec3(t, d->xmat + 9*p, MF(body_pos) + 3*b);
      xp[0] = d->xpos[3*p] + t[0]; xp[1] = d->xpos[3*p+1] + t[1]; xp[2] = d->xpos[3*p+2] + t[2];
      mulquat(xq, d->xquat + 4*p, MF(body_quat) + 4*b);
      for (int k = 0; k < jn; k++) {
        int j = ja + k;
        real *anc = d->xanchor + 3*j, *ax = d->xaxis + 3*j;
        rotvecquat(anc, MF(jnt_pos) + 3*j, xq);
        anc[0] += xp[0]; anc[1] += xp[1]; anc[2] += xp[2];
        rotvecquat(ax, MF(jnt_axis) + 3*j, xq);
        real dq = d->qpos[qadr[j]] - MF(qpos0)[qadr[j]];
        if (jtype[j] == B2_JNT_SLIDE) {
          xp[0] += ax[0]*dq; xp[1] += ax[1]*dq; xp[2] += ax[2]*dq;
        } else if (jtype[j] == B2_JNT_HINGE) {
          real ql[4], nq4[4], v[3];
          axisangle2quat(ql, MF(jnt_axis) + 3*j, dq);
          mulquat(nq4, xq, ql); memcpy(xq, nq4, sizeof(nq4));
          rotvecquat(v, MF(jnt_pos) + 3*j, xq);   /* keep the anchor fixed under the rotation */
          xp[0] = anc[0] - v[0]; xp[1] = anc[1] - v[1]; xp[2] = anc[2] - v[2];
        }
      }
    }
    normalize4(xq);
    quat2mat(d->xmat + 9*b, xq);
    real t[3], qi[4];
    mulmatvec3(t, d->xmat + 9*b, MF(body_ipos) + 3*b);
    d->xipos[3*b] = xp[0] + t[0]; d->xipos[3*b+1] = xp[1] + t[1]; d->xipos[3*b+2] = xp[2] + t[2];
    mulquat(qi, xq, MF(body_iquat) + 4*b);
    quat2mat(d->ximat + 9*b, qi);
  }
  /* world body inertial frame */
  d->xipos[0] = d->xipos[1] = d->xipos[2] = 0; quat2mat(d->ximat, d->xquat);
  for (int g = 0; g < NGEOM; g++) {
    int b = MI(geom_bodyid)[g]; real t[3], q[4];
    mulmatvec3(t, d->xmat + 9*b, MF(geom_pos) + 3*g);
    for (int k = 0; k < 3; k++) d->geom_xpos[3*g+k] = d->xpos[3*b+k] + t[k];
    mulquat(q, d->xquat + 4*b, MF(geom_quat) + 4*g); quat2mat(d->geom_xmat + 9*g, q);
  }
  for (int s = 0; s < NSITE; s++) {
    int b = MI(site_bodyid)[s]; real t[3], q[4];
    mulmatvec3(t, d->xmat + 9*b, MF(site_pos) + 3*s);
    for (int k = 0; k < 3; k++) d->site_xpos[3*s+k] = d->xpos[3*b+k] + t[k];
    mulquat(q, d->xquat + 4*b, MF(site_quat) + 4*s); quat2mat(d->site_xmat + 9*s, q);
  }
}

/* cinert = [Ixx Iyy Izz Ixy Ixz Iyz | m*off(3) | m] about the tree's com (mju_inertCom) */
static void inert_com(real *res, const real *inert, const real *mat, const real *dif, real mass) {
  real tmp[9], I[9];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) tmp[3*r+c] = mat[3*r+c] * inert[c];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
    I[3*r+c] = tmp[3*r]*mat[3*c] + tmp[3*r+1]*mat[3*c+1] + tmp[3*r+2]*mat[3*c+2];
  real d2 = dot3(dif, dif);
  res[0] = I[0] + mass*(d2 - dif[0]*dif[0]); res[1] = I[4] + mass*(d2 - dif[1]*dif[1]);
  res[2] = I[8] + mass*(d2 - dif[2]*dif[2]);
  res[3] = I[1] - mass*dif[0]*dif[1]; res[4] = I[2] - mass*dif[0]*dif[2]; res[5] = I[5] - mass*dif[1]*dif[2];
  res[6] = mass*dif[0]; res[7] = mass*dif[1]; res[8] = mass*dif[2]; res[9] = mass;
}
static void mul_inert_vec(real *res, const real *i, const real *v) {
  res[0] = i[0]*v[0] + i[3]*v[1] + i[4]*v[2] - i[8]*v[4] + i[7]*v[5];
  res[1] = i[3]*v[0] + i[1]*v[1] + i[5]*v[2] + i[8]*v[3] - i[6]*v[5];
  res[2] = i[4]*v[0] + i[5]*v[1] + i[2]*v[2] - i[7]*v[3] + i[6]*v[4];
  res[3] = i[8]*v[1] - i[7]*v[2] + i[9]*v[3];
  res[4] = i[6]*v[2] - i[8]*v[0] + i[9]*v[4];
  res[5] = i[7]*v[0] - i[6]*v[1] + i[9]*v[5];
}
static void dof_com(real *res, const real *axis, const real *offset) {
  if (offset) { res[0] = axis[0]; res[1] = axis[1]; res[2] = axis[2]; cross3(res + 3, axis, offset); }
  else { res[0] = res[1] = res[2] = 0; res[3] = axis[0]; res[4] = axis[1]; res[5] = axis[2]; }
}

static void com_pos(const RefModel *m, RefData *d) {
  const int *parent = MI(body_parentid), *rootid = MI(body_rootid);
  int nb = NBODY;
  for (int b = 0; b < nb; b++) for (int k = 0; k < 3; k++) d->subtree_com[3*b+k] = MF(body_mass)[b] * d->xipos[3*b+k];
  for (int b = nb - 1; b > 0; b--) for (int k = 0; k < 3; k++) d->subtree_com[3*parent[b]+k] += d->subtree_com[3*b+k];
  for (int b = 0; b < nb; b++) {
    real sm = MF(body_subtreemass)[b];
    if (sm < MINVAL) memcpy(d->subtree_com + 3*b, d->xipos + 3*b, 3 * sizeof(real));
    else for (int k = 0; k < 3; k++) d->subtree_com[3*b+k] /= sm;
  }
  memset(d->cinert, 0, 10 * sizeof(real));
  for (int b = 1; b < nb; b++) {
    real off[3];
    for (int k = 0; k < 3; k++) off[k] = d->xipos[3*b+k] - d->subtree_com[3*rootid[b]+k];
    inert_com(d->cinert + 10*b, MF(body_inertia) + 3*b, d->ximat + 9*b, off, MF(body_mass)[b]);
  }
  for (int j = 0; j < NJNT; j++) {
    int b = MI(jnt_bodyid)[j], da = MI(jnt_dofadr)[j];
    real off[3];
    for (int k = 0; k < 3; k++) off[k] = d->subtree_com[3*rootid[b]+k] - d->xanchor[3*j+k];
    switch (MI(jnt_type)[j]) {
      case B2_JNT_FREE: {
        real e[3];
        for (int a = 0; a < 3; a++) { e[0] = e[1] = e[2] = 0; e[a] = 1; dof_com(d->cdof + 6*(da+a), e, NULL); }
        for (int a = 0; a < 3; a++) {
          const real *xm = d->xmat + 9*b; real ax[3] = {xm[a], xm[3+a], xm[6+a]};
          dof_com(d->cdof + 6*(da+3+a), ax, off);
        }
      } break;
      case B2_JNT_SLIDE: dof_com(d->cdof + 6*da, d->xaxis + 3*j, NULL); break;
      case B2_JNT_HINGE: dof_com(d->cdof + 6*da, d->xaxis + 3*j, off); break;
    }
  }
}

/* composite rigid body inertia -> dense M (mj_crb), then Cholesky M = L L' (stands in for mj_factorM) */
static int cholesky(real *L, const real *A, int n) {
  memcpy(L, A, sizeof(real) * n * n);
  for (int j = 0; j < n; j++) {
    real s = L[j*n+j];
    for (int k = 0; k < j; k++) s -= L[j*n+k]*L[j*n+k];
    if (s < MINVAL) s = MINVAL;
    s = sqrt(s); L[j*n+j] = s;
    for (int i = j + 1; i < n; i++) {
      real t = L[i*n+j];
      for (int k = 0; k < j; k++) t -= L[i*n+k]*L[j*n+k];
      L[i*n+j] = t / s;
    }
    for (int i = 0; i < j; i++) L[i*n+j] = 0;
  }
  return 0;
}
static void chol_solve(const real *L, real *x, int n) { /* in place */
  for (int i = 0; i < n; i++) { real s = x[i]; for (int k = 0; k < i; k++) s -= L[i*n+k]*x[k]; x[i] = s / L[i*n+i]; }
  for (int i = n - 1; i >= 0; i--) { real s = x[i]; for (int k = i + 1; k < n; k++) s -= L[k*n+i]*x[k]; x[i] = s / L[i*n+i]; }
}
static void crb_factor(const RefModel *m, RefData *d) {
  int nv = NV, nb = NBODY;
  const int *parent = MI(body_parentid), *dofbody = MI(dof_bodyid), *dofparent = MI(dof_parentid);
  memcpy(d->crb, d->cinert, sizeof(real) * 10 * nb);
  for (int b = nb - 1; b > 0; b--) if (parent[b] > 0) for (int k = 0; k < 10; k++) d->crb[10*parent[b]+k] += d->crb[10*b+k];
  memset(d->M, 0, sizeof(real) * nv * nv);
  for (int i = 0; i < nv; i++) {
    real buf[6];
    mul_inert_vec(buf, d->crb + 10*dofbody[i], d->cdof + 6*i);
    for (int j = i; j >= 0; j = dofparent[j]) {
      real v = 0; for (int k = 0; k < 6; k++) v += d->cdof[6*j+k]*buf[k];
      d->M[i*nv+j] = v; d->M[j*nv+i] = v;
    }
    d->M[i*nv+i] += MF(dof_armature)[i];
  }
  cholesky(d->L, d->M, nv);
}

/* ------------------------------------------------------------------ B.4 collision (primitives) */
static void make_frame(real *f) {
  normalize3(f);
  if (norm3(f + 3) < 0.5) { f[3] = f[4] = f[5] = 0; if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1; }
  real t = dot3(f, f + 3);
  f[3] -= t*f[0]; f[4] -= t*f[1]; f[5] -= t*f[2];
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}
typedef struct { real dist, pos[3], frame[9]; } RawCon;

static int plane_sphere(RawCon *c, const real *pos1, const real *mat1, const real *pos2, real radius, real margin) {
  real n[3] = {mat1[2], mat1[5], mat1[8]}, dif[3] = {pos2[0]-pos1[0], pos2[1]-pos1[1], pos2[2]-pos1[2]};
  real cd = dot3(dif, n);
  if (cd > margin + radius) return 0;
  c->dist = cd - radius;
  real s = -c->dist * 0.5 - radius;
  for (int k = 0; k < 3; k++) { c->pos[k] = pos2[k] + n[k]*s; c->frame[k] = n[k]; c->frame[3+k] = 0; }
  return 1;
}
static int plane_capsule(RawCon *c, const real *pos1, const real *mat1, const real *pos2, const real *mat2,
                         const real *size2, real margin) {
  real ax[3] = {mat2[2], mat2[5], mat2[8]}, p[3]; int n = 0;
  for (int s = 0; s < 2; s++) {
    real sg = s ? -size2[1] : size2[1];
    for (int k = 0; k < 3; k++) p[k] = pos2[k] + ax[k]*sg;
    if (plane_sphere(c + n, pos1, mat1, p, size2[0], margin)) { memcpy(c[n].frame + 3, ax, sizeof(ax)); n++; }
  }
  return n;
}
static int plane_box(RawCon *c, const real *pos1, const real *mat1, const real *pos2, const real *mat2,
                     const real *size2, real margin) {
  real n[3] = {mat1[2], mat1[5], mat1[8]}, dif[3] = {pos2[0]-pos1[0], pos2[1]-pos1[1], pos2[2]-pos1[2]};
  real dist = dot3(dif, n); int cnt = 0;
  for (int i = 0; i < 8; i++) {
    real v[3] = {(i & 1 ? size2[0] : -size2[0]), (i & 2 ? size2[1] : -size2[1]), (i & 4 ? size2[2] : -size2[2])}, cn[3];
    mulmatvec3(cn, mat2, v);
    real ld = dot3(n, cn);
    if (dist + ld > margin || ld > 0) continue;
    c[cnt].dist = dist + ld;
    for (int k = 0; k < 3; k++) { c[cnt].pos[k] = cn[k] + pos2[k] - n[k]*c[cnt].dist*0.5; c[cnt].frame[k] = n[k]; c[cnt].frame[3+k] = 0; }
    if (++cnt >= 4) return 4;
  }
  return cnt;
}
static int plane_cylinder(RawCon *c, const real *pos1, const real *mat1, const real *pos2, const real *mat2,
                          const real *size2, real margin) {
  /* mjc_PlaneCylinder: disc-edge point nearest the plane on each cap, plus two side points on the near cap */
  real n[3] = {mat1[2], mat1[5], mat1[8]}, ax[3] = {mat2[2], mat2[5], mat2[8]};
  real dif[3] = {pos2[0]-pos1[0], pos2[1]-pos1[1], pos2[2]-pos1[2]};
  real dist0 = dot3(dif, n), prjaxis = dot3(n, ax);
  if (prjaxis > 0) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; prjaxis = -prjaxis; }
  real vec[3] = {ax[0]*prjaxis - n[0], ax[1]*prjaxis - n[1], ax[2]*prjaxis - n[2]};
  real len = norm3(vec);
  if (len < 1e-12) { vec[0] = mat2[0]*size2[0]; vec[1] = mat2[3]*size2[0]; vec[2] = mat2[6]*size2[0]; }
  else for (int k = 0; k < 3; k++) vec[k] *= size2[0] / len;
  real prjvec = dot3(vec, n);
  for (int k = 0; k < 3; k++) ax[k] *= size2[1];
  prjaxis *= size2[1];
  int cnt = 0;
  if (dist0 + prjaxis + prjvec <= margin) {
    c[cnt].dist = dist0 + prjaxis + prjvec;
    for (int k = 0; k < 3; k++) { c[cnt].pos[k] = pos2[k] + vec[k] + ax[k] - n[k]*c[cnt].dist*0.5; c[cnt].frame[k] = n[k]; c[cnt].frame[3+k] = 0; }
    cnt++;
  } else return 0;
  if (dist0 - prjaxis + prjvec <= margin) {
    c[cnt].dist = dist0 - prjaxis + prjvec;
    for (int k = 0; k < 3; k++) { c[cnt].pos[k] = pos2[k] + vec[k] - ax[k] - n[k]*c[cnt].dist*0.5; c[cnt].frame[k] = n[k]; c[cnt].frame[3+k] = 0; }
    cnt++;
  }
  real prjvec1 = -prjvec * 0.5;
  if (dist0 + prjaxis + prjvec1 <= margin) {
    real vec1[3]; cross3(vec1, vec, ax); normalize3(vec1);
    for (int k = 0; k < 3; k++) vec1[k] *= size2[0] * sqrt(3.0) * 0.5;
    for (int s = 0; s < 2; s++) {
      real sg = s ? -1.0 : 1.0;
      c[cnt].dist = dist0 + prjaxis + prjvec1;
      for (int k = 0; k < 3; k++) {
        c[cnt].pos[k] = pos2[k] + sg*vec1[k] + ax[k] - vec[k]*0.5 - n[k]*c[cnt].dist*0.5;
        c[cnt].frame[k] = n[k]; c[cnt].frame[3+k] = 0;
      }
      cnt++;
    }
  }
  return cnt;
}
static int sphere_sphere_raw(RawCon *c, const real *pos1, real r1, const real *pos2, real r2, real margin) {
  real dif[3] = {pos2[0]-pos1[0], pos2[1]-pos1[1], pos2[2]-pos1[2]};
  real cd2 = dot3(dif, dif), rs = margin + r1 + r2;
  if (cd2 > rs*rs) return 0;
  real cd = sqrt(cd2);
  c->dist = cd - r1 - r2;
  if (cd < MINVAL) { dif[0] = 1; dif[1] = dif[2] = 0; } else { dif[0] /= cd; dif[1] /= cd; dif[2] /= cd; }
  for (int k = 0; k < 3; k++) { c->frame[k] = dif[k]; c->frame[3+k] = 0; c->pos[k] = pos1[k] + dif[k]*(r1 + 0.5*c->dist); }
  return 1;
}
static int sphere_capsule(RawCon *c, const real *pos1, real r1, const real *pos2, const real *mat2, const real *size2, real margin) {
  real ax[3] = {mat2[2], mat2[5], mat2[8]}, dif[3] = {pos1[0]-pos2[0], pos1[1]-pos2[1], pos1[2]-pos2[2]};
  real x = dot3(ax, dif); if (x > size2[1]) x = size2[1]; if (x < -size2[1]) x = -size2[1];
  real p[3] = {pos2[0] + ax[0]*x, pos2[1] + ax[1]*x, pos2[2] + ax[2]*x};
  return sphere_sphere_raw(c, pos1, r1, p, size2[0], margin);
}
static int capsule_capsule(RawCon *c, const real *pos1, const real *mat1, const real *size1, const real *pos2,
                           const real *mat2, const real *size2, real margin) {
  real ax1[3] = {mat1[2], mat1[5], mat1[8]}, ax2[3] = {mat2[2], mat2[5], mat2[8]};
  real dif[3] = {pos1[0]-pos2[0], pos1[1]-pos2[1], pos1[2]-pos2[2]};
  real ma = dot3(ax1, ax1), mb = -dot3(ax1, ax2), mc = dot3(ax2, ax2);
  real u = -dot3(ax1, dif), v = dot3(ax2, dif), det = ma*mc - mb*mb;
  real p1[3], p2[3];
  if (fabs(det) >= MINVAL) {
    real x1 = (mc*u - mb*v) / det, x2 = (ma*v - mb*u) / det;
    if (x1 > size1[1]) { x1 = size1[1]; x2 = (v - mb*size1[1]) / mc; }
    else if (x1 < -size1[1]) { x1 = -size1[1]; x2 = (v + mb*size1[1]) / mc; }
    if (x2 > size2[1]) { x2 = size2[1]; x1 = (u - mb*size2[1]) / ma; if (x1 > size1[1]) x1 = size1[1]; else if (x1 < -size1[1]) x1 = -size1[1]; }
    else if (x2 < -size2[1]) { x2 = -size2[1]; x1 = (u + mb*size2[1]) / ma; if (x1 > size1[1]) x1 = size1[1]; else if (x1 < -size1[1]) x1 = -size1[1]; }
    for (int k = 0; k < 3; k++) { p1[k] = pos1[k] + ax1[k]*x1; p2[k] = pos2[k] + ax2[k]*x2; }
    return sphere_sphere_raw(c, p1, size1[0], p2, size2[0], margin);
  }
  /* parallel axes: test both ends of each segment against the other (up to 2 contacts) */
  int n = 0;
  for (int s = 0; s < 2 && n < 2; s++) {
    real x1 = s ? -size1[1] : size1[1];
    for (int k = 0; k < 3; k++) p1[k] = pos1[k] + ax1[k]*x1;
    real d2[3] = {p1[0]-pos2[0], p1[1]-pos2[1], p1[2]-pos2[2]};
    real x2 = dot3(ax2, d2); if (x2 > size2[1]) x2 = size2[1]; if (x2 < -size2[1]) x2 = -size2[1];
    for (int k = 0; k < 3; k++) p2[k] = pos2[k] + ax2[k]*x2;
    n += sphere_sphere_raw(c + n, p1, size1[0], p2, size2[0], margin);
  }
  return n;
}
static int sphere_box(RawCon *c, const real *pos1, real r1, const real *pos2, const real *mat2, const real *size2, real margin) {
  real dif[3] = {pos1[0]-pos2[0], pos1[1]-pos2[1], pos1[2]-pos2[2]}, center[3], clamped[3], deepest[3];
  /* sphere centre in box frame */
  center[0] = mat2[0]*dif[0] + mat2[3]*dif[1] + mat2[6]*dif[2];
  center[1] = mat2[1]*dif[0] + mat2[4]*dif[1] + mat2[7]*dif[2];
  center[2] = mat2[2]*dif[0] + mat2[5]*dif[1] + mat2[8]*dif[2];
  for (int k = 0; k < 3; k++) clamped[k] = center[k] > size2[k] ? size2[k] : (center[k] < -size2[k] ? -size2[k] : center[k]);
  real dv[3] = {center[0]-clamped[0], center[1]-clamped[1], center[2]-clamped[2]};
  real dist = norm3(dv), nl[3], posl[3];
  if (dist - r1 > margin) return 0;
  if (dist <= MINVAL) {
    /* centre inside the box: push out through the nearest face */
    real closest = 2 * (size2[0] + size2[1] + size2[2]); int k = 0;
    for (int i = 0; i < 6; i++) {
      real face = (i % 2 ? 1 : -1) * size2[i/2], dd = fabs(center[i/2] - face);
      if (closest > dd) { closest = dd; k = i; }
    }
    nl[0] = nl[1] = nl[2] = 0; nl[k/2] = (k % 2 ? -1 : 1);
    for (int i = 0; i < 3; i++) posl[i] = center[i] + nl[i]*(r1 - closest)*0.5;
    c->dist = -closest - r1;
  } else {
    for (int i = 0; i < 3; i++) { deepest[i] = center[i] - dv[i]/dist*r1; nl[i] = (clamped[i] - center[i]) / dist; posl[i] = 0.5*(clamped[i] + deepest[i]); }
    c->dist = dist - r1;
  }
  mulmatvec3(c->frame, mat2, nl); mulmatvec3(c->pos, mat2, posl);
  for (int k = 0; k < 3; k++) { c->pos[k] += pos2[k]; c->frame[3+k] = 0; }
  return 1;
}

/* ---- pairs beyond MuJoCo's closed-form primitives.  MuJoCo resolves cylinder-vs-{capsule,box} with its convex
 * (libccd/MPR) path and box-vs-{capsule,box} with engine_collision_box.c; neither source is available here, so the
 * functions below restate the *geometry* those routines solve (closest features, one contact for convex pairs, two for
 * a capsule lying on a face, up to eight for box faces) with deterministic tie-breaking.  Contact ordering inside a
 * pair and multi-contact selection in edge-on configurations are [EXT-unverified]; see DESIGN.md section 2. */

/* signed distance of point p (cylinder frame) to the solid cylinder (radius r, half height h); closest surface point
 * cp and outward direction n at it */
static real point_cylinder(const real *p, real r, real h, real *cp, real *n) {
  real rho = sqrt(p[0]*p[0] + p[1]*p[1]), az = fabs(p[2]), sz = p[2] < -1e-6 ? -1.0 : 1.0;   /* mid-plane ties leave upwards */
  real ux = 1, uy = 0;
  if (rho > MINVAL) { ux = p[0] / rho; uy = p[1] / rho; }
  if (az <= h && rho <= r) {                       /* inside: leave through the nearer of cap and side */
    if (h - az <= r - rho) { cp[0] = p[0]; cp[1] = p[1]; cp[2] = sz*h; n[0] = n[1] = 0; n[2] = sz; return -(h - az); }
    cp[0] = ux*r; cp[1] = uy*r; cp[2] = p[2]; n[0] = ux; n[1] = uy; n[2] = 0; return -(r - rho);
  }
  if (rho <= r) { cp[0] = p[0]; cp[1] = p[1]; cp[2] = sz*h; n[0] = n[1] = 0; n[2] = sz; return az - h; }   /* cap */
  if (az <= h) { cp[0] = ux*r; cp[1] = uy*r; cp[2] = p[2]; n[0] = ux; n[1] = uy; n[2] = 0; return rho - r; } /* side */
  cp[0] = ux*r; cp[1] = uy*r; cp[2] = sz*h;                                                                /* rim */
  real v[3] = {p[0]-cp[0], p[1]-cp[1], p[2]-cp[2]}, dd = norm3(v);
  n[0] = v[0]/dd; n[1] = v[1]/dd; n[2] = v[2]/dd;
  return dd;
}
static void to_local(real *r, const real *mat, const real *v) {   /* mat' v */
  r[0] = mat[0]*v[0] + mat[3]*v[1] + mat[6]*v[2];
  r[1] = mat[1]*v[0] + mat[4]*v[1] + mat[7]*v[2];
  r[2] = mat[2]*v[0] + mat[5]*v[1] + mat[8]*v[2];
}
/* sphere (geom1) vs cylinder (geom2): normal points from the sphere to the cylinder */
static int sphere_cylinder(RawCon *c, const real *pos1, real r1, const real *pos2, const real *mat2, const real *size2, real margin) {
  real dif[3] = {pos1[0]-pos2[0], pos1[1]-pos2[1], pos1[2]-pos2[2]}, pl[3], cp[3], n[3];
  to_local(pl, mat2, dif);
  real dist = point_cylinder(pl, size2[0], size2[1], cp, n) - r1;
  if (dist > margin) return 0;
  real posl[3] = {cp[0] + n[0]*dist*0.5, cp[1] + n[1]*dist*0.5, cp[2] + n[2]*dist*0.5}, nl[3] = {-n[0], -n[1], -n[2]};
  c->dist = dist;
  mulmatvec3(c->frame, mat2, nl); mulmatvec3(c->pos, mat2, posl);
  for (int k = 0; k < 3; k++) { c->pos[k] += pos2[k]; c->frame[3+k] = 0; }
  return 1;
}
/* ---- convex pairs: capsule-cylinder, cylinder-cylinder, cylinder-box (mjc_Convex, engine_collision_convex.c).
 * MuJoCo routes these pair types through libccd's Minkowski Portal Refinement (ccdMPRPenetration, libccd src/mpr.c;
 * the default convex path of MuJoCo 2.3 - 3.2, selectable later with the nativeccd flag off), one contact per pair
 * (multiccd is off by default).  Restated below with libccd's own tolerances and sign tests: CCD_EPS = DBL_EPSILON (MuJoCo
 * builds libccd in double), mpr_tolerance 1e-6, mpr_iterations 50 (mjOption defaults; no task overrides them).  Each geom
 * is inflated by margin/2 inside its support function (mjccd_support) and dist = margin - depth (mjc_MPRIteration).
 * mjc_fixNormal only touches spheres and ellipsoids, neither of which takes this path in the seven tasks.
 * One deviation: libccd's portal-discovery and portal-refinement loops are unbounded; they are capped at MPR_LOOP_CAP
 * passes here (and in the kernel) and report "no contact" beyond it. */
#define CCD_EPS 2.2204460492503131e-16
#define MPR_TOL 1e-6
#define MPR_MAXIT 50
#define MPR_LOOP_CAP 100
typedef struct { int type; const real *pos, *mat, *size; real margin; } CcdObj;
typedef struct { real v[3], v1[3], v2[3]; } CcdSup;
static inline int ccd_zero(real x) { return fabs(x) < CCD_EPS; }
static inline int ccd_eq(real a_, real b_) {
  real ab = fabs(a_ - b_);
  if (ab < CCD_EPS) return 1;
  real a = fabs(a_), b = fabs(b_);
  return b > a ? ab < CCD_EPS * b : ab < CCD_EPS * a;
}
static inline int ccd_vec_is_origin(const real *v) { return ccd_eq(v[0], 0) && ccd_eq(v[1], 0) && ccd_eq(v[2], 0); }
static inline real sign0(real x) { return x < 0 ? -1.0 : (x > 0 ? 1.0 : 0.0); }   /* mju_sign */
static inline void ccd_normalize(real *v) { real s = 1.0 / sqrt(dot3(v, v)); v[0] *= s; v[1] *= s; v[2] *= s; }
/* mjccd_support: farthest point of the (margin-inflated) geom along the unit direction dir, world frame */
static void ccd_support1(const CcdObj *o, const real *dir, real *res) {
  real ld[3], r[3] = {0, 0, 0};
  to_local(ld, o->mat, dir);
  switch (o->type) {
    case B2_GEOM_SPHERE: r[0] = ld[0]*o->size[0]; r[1] = ld[1]*o->size[0]; r[2] = ld[2]*o->size[0]; break;
    case B2_GEOM_CAPSULE:
      r[0] = ld[0]*o->size[0]; r[1] = ld[1]*o->size[0]; r[2] = ld[2]*o->size[0] + sign0(ld[2])*o->size[1]; break;
    case B2_GEOM_CYLINDER: {
      real t = sqrt(ld[0]*ld[0] + ld[1]*ld[1]);
      if (t > MINVAL) { r[0] = ld[0]/t*o->size[0]; r[1] = ld[1]/t*o->size[0]; }
      r[2] = sign0(ld[2])*o->size[1]; break;
    }
    case B2_GEOM_BOX: for (int k = 0; k < 3; k++) r[k] = sign0(ld[k])*o->size[k]; break;
  }
  mulmatvec3(res, o->mat, r);
  for (int k = 0; k < 3; k++) res[k] += o->pos[k] + dir[k]*o->margin;
}
static void ccd_support(const CcdObj *o1, const CcdObj *o2, const real *dir, CcdSup *s) {   /* __ccdSupport */
  real nd[3] = {-dir[0], -dir[1], -dir[2]};
  ccd_support1(o1, dir, s->v1); ccd_support1(o2, nd, s->v2);
  for (int k = 0; k < 3; k++) s->v[k] = s->v1[k] - s->v2[k];
}
static void portal_dir(const CcdSup *p, real *dir) {
  real a[3], b[3];
  for (int k = 0; k < 3; k++) { a[k] = p[2].v[k] - p[1].v[k]; b[k] = p[3].v[k] - p[1].v[k]; }
  cross3(dir, a, b); ccd_normalize(dir);
}
static int portal_reach_tolerance(const CcdSup *p, const CcdSup *v4, const real *dir) {
  real dv4 = dot3(v4->v, dir), d1 = dv4 - dot3(p[1].v, dir), d2 = dv4 - dot3(p[2].v, dir), d3 = dv4 - dot3(p[3].v, dir);
  if (d2 < d1) d1 = d2;
  if (d3 < d1) d1 = d3;
  return ccd_eq(d1, MPR_TOL) || d1 < MPR_TOL;
}
static void expand_portal(CcdSup *p, const CcdSup *v4) {
  real v4v0[3]; cross3(v4v0, v4->v, p[0].v);
  if (dot3(p[1].v, v4v0) > 0) { if (dot3(p[2].v, v4v0) > 0) p[1] = *v4; else p[3] = *v4; }
  else { if (dot3(p[3].v, v4v0) > 0) p[2] = *v4; else p[1] = *v4; }
}
/* discoverPortal: -1 no intersection, 0 portal found, 1 origin on v1, 2 origin on the v0-v1 segment */
static int discover_portal(const CcdObj *o1, const CcdObj *o2, CcdSup *p) {
  real dir[3], va[3], vb[3], dot;
  for (int k = 0; k < 3; k++) { p[0].v1[k] = o1->pos[k]; p[0].v2[k] = o2->pos[k]; p[0].v[k] = o1->pos[k] - o2->pos[k]; }
  if (ccd_vec_is_origin(p[0].v)) p[0].v[0] += CCD_EPS * 10.0;
  for (int k = 0; k < 3; k++) dir[k] = -p[0].v[k];
  ccd_normalize(dir);
  ccd_support(o1, o2, dir, &p[1]);
  dot = dot3(p[1].v, dir);
  if (ccd_zero(dot) || dot < 0) return -1;
  cross3(dir, p[0].v, p[1].v);
  if (ccd_zero(dot3(dir, dir))) return ccd_vec_is_origin(p[1].v) ? 1 : 2;
  ccd_normalize(dir);
  ccd_support(o1, o2, dir, &p[2]);
  dot = dot3(p[2].v, dir);
  if (ccd_zero(dot) || dot < 0) return -1;
  for (int k = 0; k < 3; k++) { va[k] = p[1].v[k] - p[0].v[k]; vb[k] = p[2].v[k] - p[0].v[k]; }
  cross3(dir, va, vb); ccd_normalize(dir);
  if (dot3(dir, p[0].v) > 0) { CcdSup t = p[1]; p[1] = p[2]; p[2] = t; dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2]; }
  for (int pass = 0; pass < MPR_LOOP_CAP; pass++) {
    ccd_support(o1, o2, dir, &p[3]);
    dot = dot3(p[3].v, dir);
    if (ccd_zero(dot) || dot < 0) return -1;
    int cont = 0;
    cross3(va, p[1].v, p[3].v); dot = dot3(va, p[0].v);
    if (dot < 0 && !ccd_zero(dot)) { p[2] = p[3]; cont = 1; }
    if (!cont) {
      cross3(va, p[3].v, p[2].v); dot = dot3(va, p[0].v);
      if (dot < 0 && !ccd_zero(dot)) { p[1] = p[3]; cont = 1; }
    }
    if (!cont) return 0;
    for (int k = 0; k < 3; k++) { va[k] = p[1].v[k] - p[0].v[k]; vb[k] = p[2].v[k] - p[0].v[k]; }
    cross3(dir, va, vb); ccd_normalize(dir);
  }
  return -1;
}
static int refine_portal(const CcdObj *o1, const CcdObj *o2, CcdSup *p) {
  real dir[3]; CcdSup v4;
  for (int pass = 0; pass < MPR_LOOP_CAP; pass++) {
    portal_dir(p, dir);
    real dot = dot3(dir, p[1].v);
    if (ccd_zero(dot) || dot > 0) return 0;                       /* portalEncapsulesOrigin */
    ccd_support(o1, o2, dir, &v4);
    dot = dot3(v4.v, dir);
    if (!(ccd_zero(dot) || dot > 0) || portal_reach_tolerance(p, &v4, dir)) return -1;
    expand_portal(p, &v4);
  }
  return -1;
}
static real point_seg_dist2(const real *x0, const real *b, real *wit) {   /* __ccdVec3PointSegmentDist2 with P = origin */
  real d[3] = {b[0]-x0[0], b[1]-x0[1], b[2]-x0[2]};
  real t = -dot3(x0, d) / dot3(d, d);
  if (t < 0 || ccd_zero(t)) { memcpy(wit, x0, 3*sizeof(real)); return dot3(x0, x0); }
  if (t > 1 || ccd_eq(t, 1)) { memcpy(wit, b, 3*sizeof(real)); return dot3(b, b); }
  for (int k = 0; k < 3; k++) wit[k] = d[k]*t + x0[k];
  return dot3(wit, wit);
}
static real point_tri_dist2(const real *x0, const real *B, const real *C, real *wit) {   /* ccdVec3PointTriDist2 with P = origin */
  real d1[3], d2[3];
  for (int k = 0; k < 3; k++) { d1[k] = B[k] - x0[k]; d2[k] = C[k] - x0[k]; }
  real v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(x0, d1), q = dot3(x0, d2), r = dot3(d1, d2);
  real dd = w*v - r*r, s, t;
  if (ccd_zero(dd)) s = t = -1.0;
  else { s = (q*r - w*p) / dd; t = (-s*r - q) / w; }
  if ((ccd_zero(s) || s > 0) && (ccd_eq(s, 1) || s < 1) && (ccd_zero(t) || t > 0) && (ccd_eq(t, 1) || t < 1) &&
      (ccd_eq(t + s, 1) || t + s < 1)) {
    for (int k = 0; k < 3; k++) wit[k] = x0[k] + d1[k]*s + d2[k]*t;
    return dot3(wit, wit);
  }
  real w2[3], dist = point_seg_dist2(x0, B, wit), dist2 = point_seg_dist2(x0, C, w2);
  if (dist2 < dist) { dist = dist2; memcpy(wit, w2, sizeof(w2)); }
  dist2 = point_seg_dist2(B, C, w2);
  if (dist2 < dist) { dist = dist2; memcpy(wit, w2, sizeof(w2)); }
  return dist;
}
static void find_pos(const CcdSup *p, real *pos) {
  real dir[3], vec[3], b[4], sum;
  portal_dir(p, dir);
  cross3(vec, p[1].v, p[2].v); b[0] = dot3(vec, p[3].v);
  cross3(vec, p[3].v, p[2].v); b[1] = dot3(vec, p[0].v);
  cross3(vec, p[0].v, p[1].v); b[2] = dot3(vec, p[3].v);
  cross3(vec, p[2].v, p[1].v); b[3] = dot3(vec, p[0].v);
  sum = b[0] + b[1] + b[2] + b[3];
  if (ccd_zero(sum) || sum < 0) {
    b[0] = 0;
    cross3(vec, p[2].v, p[3].v); b[1] = dot3(vec, dir);
    cross3(vec, p[3].v, p[1].v); b[2] = dot3(vec, dir);
    cross3(vec, p[1].v, p[2].v); b[3] = dot3(vec, dir);
    sum = b[1] + b[2] + b[3];
  }
  real inv = 1.0 / sum, p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) { p1[k] += p[i].v1[k]*b[i]; p2[k] += p[i].v2[k]*b[i]; }
  for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p1[k]*inv + p2[k]*inv);
}
/* ccdMPRPenetration + mjc_MPRIteration: 0 or 1 contact, normal from geom 1 to geom 2 */
static int mpr_convex(RawCon *c, int t1, const real *pos1, const real *mat1, const real *size1, int t2, const real *pos2,
                      const real *mat2, const real *size2, real margin) {
  CcdObj o1 = {t1, pos1, mat1, size1, 0.5*margin}, o2 = {t2, pos2, mat2, size2, 0.5*margin};
  CcdSup p[4]; real depth, dir[3], pos[3];
  int res = discover_portal(&o1, &o2, p);
  if (res < 0) return 0;
  if (res == 1) return 0;            /* findPenetrTouch: depth 0, zero direction -> mjc_MPRIteration discards it */
  if (res == 2) {                    /* findPenetrSegment */
    for (int k = 0; k < 3; k++) { pos[k] = 0.5*(p[1].v1[k] + p[1].v2[k]); dir[k] = p[1].v[k]; }
    depth = sqrt(dot3(dir, dir)); ccd_normalize(dir);
  } else {
    if (refine_portal(&o1, &o2, p) < 0) return 0;
    CcdSup v4; real pd[3];
    for (int it = 0; ; it++) {       /* findPenetr */
      portal_dir(p, pd);
      ccd_support(&o1, &o2, pd, &v4);
      if (portal_reach_tolerance(p, &v4, pd) || it > MPR_MAXIT) {
        depth = sqrt(point_tri_dist2(p[1].v, p[2].v, p[3].v, dir));
        if (ccd_zero(depth)) return 0;                             /* zero direction: discarded */
        ccd_normalize(dir);
        find_pos(p, pos);
        break;
      }
      expand_portal(p, &v4);
    }
  }
  if (ccd_vec_is_origin(dir)) return 0;
  if (!(depth == depth) || !(dir[0] == dir[0]) || !(pos[0] == pos[0]) || !(pos[1] == pos[1]) || !(pos[2] == pos[2])) return 0;   /* collapsed portal */
  c->dist = margin - depth;
  for (int k = 0; k < 3; k++) { c->pos[k] = pos[k]; c->frame[k] = dir[k]; c->frame[3+k] = 0; }
  normalize3(c->frame);
  return 1;
}

/* squared distance from the point c + a t (box frame) to the box with half sizes s */
static real seg_box_d2(const real *c, const real *a, const real *s, real t) {
  real f = 0;
  for (int k = 0; k < 3; k++) { real p = fabs(c[k] + a[k]*t) - s[k]; if (p > 0) f += p*p; }
  return f;
}
/* capsule (geom1) vs box (geom2).  t* = segment parameter closest to the box (exact minimisation of the piecewise
 * quadratic distance, smallest t on ties); contact 1 = sphere-box at t*.  If that contact is on a face (exactly one
 * coordinate outside the box), a second sphere-box test is made at the far end of the part of the segment that lies
 * over the same face; the two contacts are emitted in ascending t. */
static int capsule_box(RawCon *c, const real *pos1, const real *mat1, const real *size1, const real *pos2,
                       const real *mat2, const real *size2, real margin) {
  real ax[3] = {mat1[2], mat1[5], mat1[8]}, dif[3] = {pos1[0]-pos2[0], pos1[1]-pos2[1], pos1[2]-pos2[2]}, cl[3], al[3];
  to_local(cl, mat2, dif); to_local(al, mat2, ax);
  real l = size1[1], r = size1[0];
  {                                                  /* axis passes through the box: one contact at the chord point nearest the capsule centre */
    real te = -l, tx = l; int hit = 1;
    for (int k = 0; k < 3; k++) {
      if (fabs(al[k]) <= 1e-9) { if (fabs(cl[k]) > size2[k]) hit = 0; continue; }
      real u = (-size2[k] - cl[k]) / al[k], v = (size2[k] - cl[k]) / al[k];
      if (u > v) { real w = u; u = v; v = w; }
      if (u > te) te = u; if (v < tx) tx = v;
    }
    if (hit && te <= tx) {
      real tm = te > 0 ? te : (tx < 0 ? tx : 0), pm[3] = {pos1[0] + ax[0]*tm, pos1[1] + ax[1]*tm, pos1[2] + ax[2]*tm};
      return sphere_box(c, pm, r, pos2, mat2, size2, margin);
    }
  }
  real cand[24]; int nc = 0;
  cand[nc++] = -l; cand[nc++] = l;
  for (int k = 0; k < 3; k++) if (fabs(al[k]) > 1e-9) for (int sg = -1; sg <= 1; sg += 2) {
    real t = (sg*size2[k] - cl[k]) / al[k];
    if (t > -l && t < l) cand[nc++] = t;
  }
  for (int i = 1; i < nc; i++) { real v = cand[i]; int j = i - 1; while (j >= 0 && cand[j] > v) { cand[j+1] = cand[j]; j--; } cand[j+1] = v; }
  int nb = nc;
  for (int i = 0; i + 1 < nb; i++) {                 /* stationary point of the quadratic on each interval */
    real ta = cand[i], tb = cand[i+1], tm = 0.5*(ta + tb), num = 0, den = 0;
    for (int k = 0; k < 3; k++) {
      real p = cl[k] + al[k]*tm;
      if (p > size2[k]) { num += al[k]*(cl[k] - size2[k]); den += al[k]*al[k]; }
      else if (p < -size2[k]) { num += al[k]*(cl[k] + size2[k]); den += al[k]*al[k]; }
    }
    if (den > 1e-12) { real t = -num / den; if (t > ta && t < tb) cand[nc++] = t; }
  }
  real tbest = cand[0], fbest = seg_box_d2(cl, al, size2, cand[0]);
  for (int i = 1; i < nc; i++) {
    real f = seg_box_d2(cl, al, size2, cand[i]);
    if (f < fbest || (f == fbest && cand[i] < tbest)) { fbest = f; tbest = cand[i]; }
  }
  real p1[3] = {pos1[0] + ax[0]*tbest, pos1[1] + ax[1]*tbest, pos1[2] + ax[2]*tbest};
  RawCon first; int n1 = sphere_box(&first, p1, r, pos2, mat2, size2, margin);
  if (!n1) return 0;
  /* face contact? */
  int nout = 0, kf = -1;
  for (int k = 0; k < 3; k++) if (fabs(cl[k] + al[k]*tbest) > size2[k] + 1e-6 + 1e-5*size2[k]) { nout++; kf = k; }
  int n = 0; real t2 = tbest; int have2 = 0; RawCon second;
  if (nout == 1) {
    real ta = -l, tb = l; int ok = 1;
    for (int k = 0; k < 3; k++) if (k != kf) {
      if (fabs(al[k]) <= 1e-9) { if (fabs(cl[k]) > size2[k]) ok = 0; continue; }
      real u = (-size2[k] - cl[k]) / al[k], v = (size2[k] - cl[k]) / al[k];
      if (u > v) { real w = u; u = v; v = w; }
      if (u > ta) ta = u; if (v < tb) tb = v;
    }
    if (ok && tb >= ta) {
      t2 = (tbest - ta > tb - tbest) ? ta : tb;
      if (fabs(t2 - tbest) > 1e-3*l) {
        real p2[3] = {pos1[0] + ax[0]*t2, pos1[1] + ax[1]*t2, pos1[2] + ax[2]*t2};
        have2 = sphere_box(&second, p2, r, pos2, mat2, size2, margin);
      }
    }
  }
  if (have2 && t2 < tbest) { c[n++] = second; c[n++] = first; }
  else { c[n++] = first; if (have2) c[n++] = second; }
  return n;
}

/* box (geom1) vs box (geom2): separating-axis test over the 15 axes; a face axis wins unless an edge axis is
 * better by a clear margin.  Face case: the incident face is clipped against the side planes of the reference face
 * (Sutherland-Hodgman, <= 8 points), points farther than `margin` from the reference face are dropped.  Edge case: one
 * contact between the closest points of the two edges.  The frame normal points from box 1 to box 2. */
static int box_box(RawCon *c, const real *pos1, const real *mat1, const real *size1, const real *pos2,
                   const real *mat2, const real *size2, real margin) {
  real A[3][3], B[3][3], p[3] = {pos2[0]-pos1[0], pos2[1]-pos1[1], pos2[2]-pos1[2]};
  for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) { A[i][k] = mat1[3*k+i]; B[i][k] = mat2[3*k+i]; }   /* axes as rows */
  real R[3][3], Q[3][3], pa[3], pb[3];
  for (int i = 0; i < 3; i++) { pa[i] = dot3(p, A[i]); pb[i] = dot3(p, B[i]); for (int j = 0; j < 3; j++) { R[i][j] = dot3(A[i], B[j]); Q[i][j] = fabs(R[i][j]); } }
  real best = -1e300; int code = -1; real bsign = 1;
  for (int i = 0; i < 3; i++) {                    /* faces of box 1 */
    real s = fabs(pa[i]) - (size1[i] + size2[0]*Q[i][0] + size2[1]*Q[i][1] + size2[2]*Q[i][2]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = i; bsign = pa[i] < 0 ? -1 : 1; }
  }
  for (int j = 0; j < 3; j++) {                    /* faces of box 2 */
    real s = fabs(pb[j]) - (size2[j] + size1[0]*Q[0][j] + size1[1]*Q[1][j] + size1[2]*Q[2][j]);
    if (s > margin) return 0;
    if (s > best) { best = s; code = 3 + j; bsign = pb[j] < 0 ? -1 : 1; }
  }
  real ebest = -1e300; int ecode = -1; real en[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    real L[3]; cross3(L, A[i], B[j]);
    real len = norm3(L);
    if (len < 1e-6) continue;
    for (int k = 0; k < 3; k++) L[k] /= len;
    real ra = 0, rb = 0;
    for (int k = 0; k < 3; k++) { ra += size1[k]*fabs(dot3(A[k], L)); rb += size2[k]*fabs(dot3(B[k], L)); }
    real pl = dot3(p, L), s = fabs(pl) - (ra + rb);
    if (s > margin) return 0;
    if (s > ebest) { ebest = s; ecode = 3*i + j; real sg = pl < 0 ? -1 : 1; en[0] = L[0]*sg; en[1] = L[1]*sg; en[2] = L[2]*sg; }
  }
  if (ecode >= 0 && ebest > best + 1e-4 + 0.02*fabs(best)) {
    /* edge-edge: edge of box 1 along A[i], of box 2 along B[j], both through the corner facing the other box */
    int i = ecode / 3, j = ecode % 3; real ea[3], eb[3];
    for (int k = 0; k < 3; k++) { ea[k] = pos1[k]; eb[k] = pos2[k]; }
    for (int a = 0; a < 3; a++) if (a != i) { real sg = dot3(en, A[a]) > 0 ? 1 : -1; for (int k = 0; k < 3; k++) ea[k] += sg*size1[a]*A[a][k]; }
    for (int b = 0; b < 3; b++) if (b != j) { real sg = dot3(en, B[b]) > 0 ? -1 : 1; for (int k = 0; k < 3; k++) eb[k] += sg*size2[b]*B[b][k]; }
    real w[3] = {eb[0]-ea[0], eb[1]-ea[1], eb[2]-ea[2]}, uaub = dot3(A[i], B[j]), q1 = dot3(A[i], w), q2 = -dot3(B[j], w);
    real dd = 1 - uaub*uaub, alpha = 0, beta = 0;
    if (dd > 1e-12) { alpha = (q1 + uaub*q2) / dd; beta = (uaub*q1 + q2) / dd; }
    if (alpha > size1[i]) alpha = size1[i]; if (alpha < -size1[i]) alpha = -size1[i];
    if (beta > size2[j]) beta = size2[j]; if (beta < -size2[j]) beta = -size2[j];
    for (int k = 0; k < 3; k++) {
      real x1 = ea[k] + alpha*A[i][k], x2 = eb[k] + beta*B[j][k];
      c->pos[k] = 0.5*(x1 + x2); c->frame[k] = en[k]; c->frame[3+k] = 0;
    }
    c->dist = ebest;
    return 1;
  }
  /* face case */
  const real (*Rf)[3], (*If)[3]; const real *rs, *is, *rp, *ip; int fi; real nout[3];
  if (code < 3) { Rf = A; If = B; rs = size1; is = size2; rp = pos1; ip = pos2; fi = code; for (int k = 0; k < 3; k++) nout[k] = bsign*A[fi][k]; }
  else { Rf = B; If = A; rs = size2; is = size1; rp = pos2; ip = pos1; fi = code - 3; for (int k = 0; k < 3; k++) nout[k] = -bsign*B[fi][k]; }
  /* incident face: the face of the other box most anti-parallel to nout */
  int ii = 0; real mn = 1e300, isg = 1;
  for (int k = 0; k < 3; k++) { real dn = dot3(If[k], nout); if (-fabs(dn) < mn) { mn = -fabs(dn); ii = k; isg = dn > 0 ? -1 : 1; } }
  int i1 = (ii + 1) % 3, i2 = (ii + 2) % 3, r1 = (fi + 1) % 3, r2 = (fi + 2) % 3;
  real poly[16][3], tmp[16][3]; int np = 4;
  static const int sx[4] = {1, -1, -1, 1}, sy[4] = {1, 1, -1, -1};
  for (int v = 0; v < 4; v++) for (int k = 0; k < 3; k++)
    poly[v][k] = ip[k] + isg*is[ii]*If[ii][k] + sx[v]*is[i1]*If[i1][k] + sy[v]*is[i2]*If[i2][k] - rp[k];
  /* clip against the four side planes  +-Rf[r1] . x <= rs[r1],  +-Rf[r2] . x <= rs[r2]  (x relative to the reference centre) */
  for (int pl = 0; pl < 4 && np > 0; pl++) {
    const real *axv = Rf[pl < 2 ? r1 : r2]; real sg = (pl & 1) ? -1 : 1, lim = rs[pl < 2 ? r1 : r2]; int nn = 0;
    for (int v = 0; v < np; v++) {
      const real *a = poly[v], *b = poly[(v + 1) % np];
      real da = sg*dot3(axv, a) - lim, db = sg*dot3(axv, b) - lim;
      if (da <= 0) { memcpy(tmp[nn++], a, sizeof(real)*3); }
      if ((da < 0 && db > 0) || (da > 0 && db < 0)) { real t = da / (da - db); for (int k = 0; k < 3; k++) tmp[nn][k] = a[k] + t*(b[k] - a[k]); nn++; }
    }
    np = nn > 8 ? 8 : nn;
    memcpy(poly, tmp, sizeof(real)*3*np);
  }
  int cnt = 0; real fsign = (code < 3) ? 1.0 : -1.0;   /* nout is the outward normal of the reference face */
  for (int v = 0; v < np && cnt < 8; v++) {
    real depth = dot3(poly[v], nout) - rs[fi];
    if (depth > margin) continue;
    c[cnt].dist = depth;
    for (int k = 0; k < 3; k++) { c[cnt].pos[k] = poly[v][k] + rp[k] - nout[k]*depth*0.5; c[cnt].frame[k] = fsign*nout[k]; c[cnt].frame[3+k] = 0; }
    cnt++;
  }
  return cnt;
}

/* direct access to the narrow phase for differential tests: types/poses/sizes in, up to 8 contacts of
 * [dist, pos(3), normal(3), tangent hint(3)] out; returns the count (-1: pair type outside the subset) */
int ref_collide_raw(int t1, int t2, const double *p1, const double *m1, const double *s1, const double *p2,
                    const double *m2, const double *s2, double margin, double *out80) {
  RawCon raw[8]; int n = -1;
  memset(raw, 0, sizeof(raw));
  if (t1 == B2_GEOM_PLANE) {
    if (t2 == B2_GEOM_SPHERE) n = plane_sphere(raw, p1, m1, p2, s2[0], margin);
    else if (t2 == B2_GEOM_CAPSULE) n = plane_capsule(raw, p1, m1, p2, m2, s2, margin);
    else if (t2 == B2_GEOM_BOX) n = plane_box(raw, p1, m1, p2, m2, s2, margin);
    else if (t2 == B2_GEOM_CYLINDER) n = plane_cylinder(raw, p1, m1, p2, m2, s2, margin);
  } else if (t1 == B2_GEOM_SPHERE) {
    if (t2 == B2_GEOM_SPHERE) n = sphere_sphere_raw(raw, p1, s1[0], p2, s2[0], margin);
    else if (t2 == B2_GEOM_CAPSULE) n = sphere_capsule(raw, p1, s1[0], p2, m2, s2, margin);
    else if (t2 == B2_GEOM_BOX) n = sphere_box(raw, p1, s1[0], p2, m2, s2, margin);
    else if (t2 == B2_GEOM_CYLINDER) n = sphere_cylinder(raw, p1, s1[0], p2, m2, s2, margin);
  } else if (t1 == B2_GEOM_CAPSULE) {
    if (t2 == B2_GEOM_CAPSULE) n = capsule_capsule(raw, p1, m1, s1, p2, m2, s2, margin);
    else if (t2 == B2_GEOM_CYLINDER) n = mpr_convex(raw, t1, p1, m1, s1, t2, p2, m2, s2, margin);
    else if (t2 == B2_GEOM_BOX) n = capsule_box(raw, p1, m1, s1, p2, m2, s2, margin);
  } else if (t1 == B2_GEOM_CYLINDER && t2 == B2_GEOM_BOX) n = mpr_convex(raw, t1, p1, m1, s1, t2, p2, m2, s2, margin);
  else if (t1 == B2_GEOM_CYLINDER && t2 == B2_GEOM_CYLINDER) n = mpr_convex(raw, t1, p1, m1, s1, t2, p2, m2, s2, margin);
  else if (t1 == B2_GEOM_BOX && t2 == B2_GEOM_BOX) n = box_box(raw, p1, m1, s1, p2, m2, s2, margin);
  for (int k = 0; k < n && k < 8; k++) { out80[10*k] = raw[k].dist; memcpy(out80 + 10*k + 1, raw[k].pos, 3*sizeof(real)); memcpy(out80 + 10*k + 4, raw[k].frame, 6*sizeof(real)); }
  return n;
}

/* the MPR path on any pair of sphere / capsule / cylinder / box (known-answer tests against the closed-form pairs) */
int ref_mpr_raw(int t1, int t2, const double *p1, const double *m1, const double *s1, const double *p2,
                const double *m2, const double *s2, double margin, double *out10) {
  RawCon raw; memset(&raw, 0, sizeof(raw));
  int n = mpr_convex(&raw, t1, p1, m1, s1, t2, p2, m2, s2, margin);
  if (n) { out10[0] = raw.dist; memcpy(out10 + 1, raw.pos, 3*sizeof(real)); memcpy(out10 + 4, raw.frame, 6*sizeof(real)); }
  return n;
}

static int collide_pair(const RefModel *m, const RefData *d, int g1, int g2, real margin, RawCon *out) {
  int t1 = MI(geom_type)[g1], t2 = MI(geom_type)[g2];
  const real *p1 = d->geom_xpos + 3*g1, *p2 = d->geom_xpos + 3*g2, *m1 = d->geom_xmat + 9*g1, *m2 = d->geom_xmat + 9*g2;
  const real *s1 = MF(geom_size) + 3*g1, *s2 = MF(geom_size) + 3*g2;
  if (t1 == B2_GEOM_PLANE) {
    switch (t2) {
      case B2_GEOM_SPHERE: return plane_sphere(out, p1, m1, p2, s2[0], margin);
      case B2_GEOM_CAPSULE: return plane_capsule(out, p1, m1, p2, m2, s2, margin);
      case B2_GEOM_BOX: return plane_box(out, p1, m1, p2, m2, s2, margin);
      case B2_GEOM_CYLINDER: return plane_cylinder(out, p1, m1, p2, m2, s2, margin);
    }
  } else if (t1 == B2_GEOM_SPHERE) {
    switch (t2) {
      case B2_GEOM_SPHERE: return sphere_sphere_raw(out, p1, s1[0], p2, s2[0], margin);
      case B2_GEOM_CAPSULE: return sphere_capsule(out, p1, s1[0], p2, m2, s2, margin);
      case B2_GEOM_BOX: return sphere_box(out, p1, s1[0], p2, m2, s2, margin);
      case B2_GEOM_CYLINDER: return sphere_cylinder(out, p1, s1[0], p2, m2, s2, margin);
    }
  } else if (t1 == B2_GEOM_CAPSULE) {
    switch (t2) {
      case B2_GEOM_CAPSULE: return capsule_capsule(out, p1, m1, s1, p2, m2, s2, margin);
      case B2_GEOM_CYLINDER: return mpr_convex(out, t1, p1, m1, s1, t2, p2, m2, s2, margin);
      case B2_GEOM_BOX: return capsule_box(out, p1, m1, s1, p2, m2, s2, margin);
    }
  } else if (t1 == B2_GEOM_CYLINDER && t2 == B2_GEOM_BOX) {
    return mpr_convex(out, t1, p1, m1, s1, t2, p2, m2, s2, margin);
  } else if (t1 == B2_GEOM_CYLINDER && t2 == B2_GEOM_CYLINDER) {
    return mpr_convex(out, t1, p1, m1, s1, t2, p2, m2, s2, margin);
  } else if (t1 == B2_GEOM_BOX && t2 == B2_GEOM_BOX) {
    return box_box(out, p1, m1, s1, p2, m2, s2, margin);
  }
  return -1; /* pair type outside the restated subset */
}

static int collision(const RefModel *m, RefData *d) {
  d->ncon = 0; d->ncon_dropped = 0;
  int unsupported = 0;
  for (int p = 0; p < NPAIR; p++) {
    int g1 = MI(pair_g1)[p], g2 = MI(pair_g2)[p];
    real margin = MF(pair_margin)[p];
    /* bounding-sphere / plane-distance cull (cannot change results: only prunes) */
    real rb1 = MF(geom_rbound)[g1], rb2 = MF(geom_rbound)[g2];
    const real *p1 = d->geom_xpos + 3*g1, *p2 = d->geom_xpos + 3*g2;
    if (MI(geom_type)[g1] == B2_GEOM_PLANE) {
      const real *m1 = d->geom_xmat + 9*g1; real n[3] = {m1[2], m1[5], m1[8]};
      real dif[3] = {p2[0]-p1[0], p2[1]-p1[1], p2[2]-p1[2]};
      if (dot3(dif, n) > rb2 + margin) continue;
    } else {
      real dif[3] = {p2[0]-p1[0], p2[1]-p1[1], p2[2]-p1[2]}, bound = rb1 + rb2 + margin;
      if (dot3(dif, dif) > bound*bound) continue;
    }
    RawCon raw[8];
    int n = collide_pair(m, d, g1, g2, margin, raw);
    if (n < 0) { unsupported++; continue; }
    for (int k = 0; k < n; k++) {
      if (d->ncon >= d->maxcon) { d->ncon_dropped++; continue; }
      RefContact *c = d->contact + d->ncon++;
      c->dist = raw[k].dist; memcpy(c->pos, raw[k].pos, sizeof(c->pos)); memcpy(c->frame, raw[k].frame, 6 * sizeof(real));
      make_frame(c->frame);
      c->includemargin = margin - MF(pair_gap)[p];
      memcpy(c->friction, MF(pair_friction) + 5*p, 5 * sizeof(real));
      memcpy(c->solref, MF(pair_solref) + 2*p, 2 * sizeof(real));
      memcpy(c->solimp, MF(pair_solimp) + 5*p, 5 * sizeof(real));
      c->dim = MI(pair_condim)[p]; c->geom1 = g1; c->geom2 = g2; c->pair = p; c->efc_address = -1;
    }
  }
  return unsupported;
}

/* ------------------------------------------------------------------ B.5 constraints */
/* translational Jacobian of `point` attached to `body`, dense 3 x nv (mj_jac) */
static void jac_point(const RefModel *m, const RefData *d, real *jacp, real *jacr, const real *point, int body) {
  int nv = NV;
  memset(jacp, 0, sizeof(real) * 3 * nv); if (jacr) memset(jacr, 0, sizeof(real) * 3 * nv);
  int i = MI(body_lastdof)[body];
  real off[3]; const real *com = d->subtree_com + 3*MI(body_rootid)[body];
  for (int k = 0; k < 3; k++) off[k] = point[k] - com[k];
  for (; i >= 0; i = MI(dof_parentid)[i]) {
    const real *c = d->cdof + 6*i; real t[3];
    cross3(t, c, off);
    for (int k = 0; k < 3; k++) { jacp[k*nv+i] = c[3+k] + t[k]; if (jacr) jacr[k*nv+i] = c[k]; }
  }
}
static void ensure_efc(const RefModel *m, RefData *d, int nefc) {
  if (nefc <= d->arcap) return;
  int cap = nefc + 64;
  free(d->efc_J); free(d->efc_AR);
  d->efc_J = zalloc((size_t)cap * NV); d->efc_AR = zalloc((size_t)cap * cap); d->arcap = cap;
}
static real get_impedance(const real *solimp_in, real pos, real margin) {
  real s[5]; memcpy(s, solimp_in, sizeof(s));
  if (s[0] < MINIMP) s[0] = MINIMP; if (s[0] > MAXIMP) s[0] = MAXIMP;
  if (s[1] < MINIMP) s[1] = MINIMP; if (s[1] > MAXIMP) s[1] = MAXIMP;
  if (s[2] < 0) s[2] = 0;
  if (s[3] < MINIMP) s[3] = MINIMP; if (s[3] > MAXIMP) s[3] = MAXIMP;
  if (s[4] < 1) s[4] = 1;
  if (s[0] == s[1] || s[2] <= MINVAL) return 0.5 * (s[0] + s[1]);
  real x = fabs(pos - margin) / s[2];
  if (x >= 1) return s[1];
  if (x <= 0) return s[0];
  real y;
  if (s[4] == 1) y = x;
  else if (x <= s[3]) { real a = 1.0 / pow(s[3], s[4] - 1); y = a * pow(x, s[4]); }
  else { real b = 1.0 / pow(1 - s[3], s[4] - 1); y = 1 - b * pow(1 - x, s[4]); }
  return s[0] + y * (s[1] - s[0]);
}

static void make_constraint(const RefModel *m, RefData *d) {
  int nv = NV, nlim = m->d[B2D_nlimited];
  /* count rows first so the dense buffers can be sized */
  int nrow = 0;
  for (int k = 0; k < nlim; k++) {
    int j = MI(limited_jnt)[k]; real q = d->qpos[MI(jnt_qposadr)[j]], mg = MF(jnt_margin)[j];
    if (q - MF(jnt_range)[2*j] < mg) nrow++;
    if (MF(jnt_range)[2*j+1] - q < mg) nrow++;
  }
  for (int c = 0; c < d->ncon; c++) nrow += d->contact[c].dim == 1 ? 1 : 2 * (d->contact[c].dim - 1);
  ensure_efc(m, d, nrow);
  if (nrow > d->maxefc) nrow = d->maxefc;
  int r = 0;
  /* joint limits, lower side then upper side (mj_instantiateLimit) */
  for (int k = 0; k < nlim; k++) {
    int j = MI(limited_jnt)[k], da = MI(jnt_dofadr)[j]; real q = d->qpos[MI(jnt_qposadr)[j]], mg = MF(jnt_margin)[j];
    for (int side = -1; side <= 1; side += 2) {
      real dist = side * (MF(jnt_range)[2*j + (side + 1) / 2] - q);
      if (dist < mg && r < nrow) {
        memset(d->efc_J + (size_t)r*nv, 0, sizeof(real) * nv);
        d->efc_J[(size_t)r*nv + da] = -(real)side;
        d->efc_pos[r] = dist; d->efc_margin[r] = mg; d->efc_type[r] = CNSTR_LIMIT; d->efc_id[r] = j; r++;
      }
    }
  }
  /* pyramidal contacts (mj_instantiateContact) */
  real *jp1 = d->scratch, *jp2 = jp1 + 3*nv, *jr1 = jp2 + 3*nv, *jr2 = jr1 + 3*nv;
  for (int c = 0; c < d->ncon; c++) {
    RefContact *con = d->contact + c;
    int nr = con->dim == 1 ? 1 : 2 * (con->dim - 1);
    if (r + nr > nrow) { con->efc_address = -1; continue; }
    con->efc_address = r;
    int b1 = MI(geom_bodyid)[con->geom1], b2 = MI(geom_bodyid)[con->geom2];
    jac_point(m, d, jp1, jr1, con->pos, b1); jac_point(m, d, jp2, jr2, con->pos, b2);
    /* rows of the contact frame applied to (J2 - J1): [normal, t1, t2, (rot) normal, t1, t2] */
    for (int row = 0; row < nr; row++) {
      real *J = d->efc_J + (size_t)(r + row)*nv;
      int kdir = con->dim == 1 ? 0 : 1 + row / 2; real sgn = (row % 2) ? -1.0 : 1.0;
      real mu = con->dim == 1 ? 0 : con->friction[kdir - 1];
      for (int i = 0; i < nv; i++) {
        real jn = 0, jt = 0;
        for (int a = 0; a < 3; a++) jn += con->frame[a] * (jp2[a*nv+i] - jp1[a*nv+i]);
        if (con->dim > 1) {
          if (kdir < 3) for (int a = 0; a < 3; a++) jt += con->frame[3*kdir + a] * (jp2[a*nv+i] - jp1[a*nv+i]);
          else for (int a = 0; a < 3; a++) jt += con->frame[3*(kdir-3) + a] * (jr2[a*nv+i] - jr1[a*nv+i]);
        }
        J[i] = jn + sgn * mu * jt;
      }
      d->efc_pos[r+row] = con->dist; d->efc_margin[r+row] = con->includemargin;
      d->efc_type[r+row] = con->dim == 1 ? CNSTR_CONTACT_FRICTIONLESS : CNSTR_CONTACT_PYRAMIDAL; d->efc_id[r+row] = c;
    }
    r += nr;
  }
  d->nefc = r;
  /* diagApprox, impedance, R, D, KBIP (mj_diagApprox + mj_makeImpedance) */
  for (int i = 0; i < d->nefc; i++) {
    const real *solref, *solimp; real da;
    if (d->efc_type[i] == CNSTR_LIMIT) {
      int j = d->efc_id[i];
      solref = MF(jnt_solref) + 2*j; solimp = MF(jnt_solimp) + 5*j; da = MF(dof_invweight0)[MI(jnt_dofadr)[j]];
    } else {
      RefContact *con = d->contact + d->efc_id[i];
      int b1 = MI(geom_bodyid)[con->geom1], b2 = MI(geom_bodyid)[con->geom2];
      real tran = MF(body_invweight0)[2*b1] + MF(body_invweight0)[2*b2];
      real rot = MF(body_invweight0)[2*b1+1] + MF(body_invweight0)[2*b2+1];
      solref = con->solref; solimp = con->solimp;
      if (con->dim == 1) da = tran;
      else {
        int k = (i - con->efc_address) / 2; real fr = con->friction[k];
        da = tran + fr*fr*(k < 2 ? tran : rot);
      }
    }
    if (da < MINVAL) da = MINVAL;
    d->efc_diagApprox[i] = da;
    real imp = get_impedance(solimp, d->efc_pos[i], d->efc_margin[i]);
    real R = (1 - imp) * da / imp; if (R < MINVAL) R = MINVAL;
    d->efc_R[i] = R;
    real dmax = solimp[1]; if (dmax < MINIMP) dmax = MINIMP; if (dmax > MAXIMP) dmax = MAXIMP;
    real K, B;
    if (solref[0] > 0) {
      real tc = solref[0], dr = solref[1];
      if (tc < 2 * m->timestep) tc = 2 * m->timestep;           /* refsafe */
      real den = dmax*dmax*tc*tc*dr*dr; K = 1.0 / (den > MINVAL ? den : MINVAL);
      den = dmax*tc; B = 2.0 / (den > MINVAL ? den : MINVAL);
    } else { K = -solref[0] / (dmax*dmax); B = -solref[1] / dmax; }
    d->efc_KBIP[4*i] = K; d->efc_KBIP[4*i+1] = B; d->efc_KBIP[4*i+2] = imp; d->efc_KBIP[4*i+3] = 0;
  }
  /* pyramidal contacts share one regulariser: R = 2 mu^2 R(first row) */
  for (int c = 0; c < d->ncon; c++) {
    RefContact *con = d->contact + c;
    if (con->efc_address < 0 || con->dim == 1) continue;
    int a = con->efc_address, nr = 2 * (con->dim - 1);
    real mu = con->friction[0] * sqrt(1.0 / m->impratio);
    real Rpy = 2 * mu * mu * d->efc_R[a]; if (Rpy < MINVAL) Rpy = MINVAL;
    for (int k = 0; k < nr; k++) d->efc_R[a+k] = Rpy;
  }
  for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1.0 / d->efc_R[i];
}

/* A_R = J M^-1 J' + diag(R)  (mj_projectConstraint) */
static void project_constraint(const RefModel *m, RefData *d) {
  int nv = NV, ne = d->nefc;
  if (!ne) return;
  real *B = (real *)malloc(sizeof(real) * (size_t)ne * nv);
  for (int i = 0; i < ne; i++) { memcpy(B + (size_t)i*nv, d->efc_J + (size_t)i*nv, sizeof(real) * nv); chol_solve(d->L, B + (size_t)i*nv, nv); }
  for (int i = 0; i < ne; i++) for (int j = 0; j <= i; j++) {
    real s = 0; const real *Ji = d->efc_J + (size_t)i*nv, *Bj = B + (size_t)j*nv;
    for (int k = 0; k < nv; k++) s += Ji[k]*Bj[k];
    d->efc_AR[(size_t)i*ne+j] = s; d->efc_AR[(size_t)j*ne+i] = s;
  }
  for (int i = 0; i < ne; i++) d->efc_AR[(size_t)i*ne+i] += d->efc_R[i];
  free(B);
}

static void fwd_position(const RefModel *m, RefData *d) {
  kinematics(m, d); com_pos(m, d); crb_factor(m, d); collision(m, d); make_constraint(m, d);
  if (m->d[B2D_solver] == B2_SOLVER_PGS) project_constraint(m, d);
}

/* ------------------------------------------------------------------ B.2 velocity stage */
static void cross_motion(real *res, const real *vel, const real *v) {
  real t[3];
  cross3(res, vel, v);
  cross3(res + 3, vel, v + 3); cross3(t, vel + 3, v);
  res[3] += t[0]; res[4] += t[1]; res[5] += t[2];
}
static void cross_force(real *res, const real *vel, const real *f) {
  real t[3];
  cross3(res, vel, f); cross3(t, vel + 3, f + 3);
  res[0] += t[0]; res[1] += t[1]; res[2] += t[2];
  cross3(res + 3, vel, f + 3);
}
static void com_vel(const RefModel *m, RefData *d) {
  memset(d->cvel, 0, 6 * sizeof(real));
  for (int b = 1; b < NBODY; b++) {
    real cv[6]; memcpy(cv, d->cvel + 6*MI(body_parentid)[b], sizeof(cv));
    int ja = MI(body_jntadr)[b], jn = MI(body_jntnum)[b];
    for (int k = 0; k < jn; k++) {
      int j = ja + k, da = MI(jnt_dofadr)[j];
      if (MI(jnt_type)[j] == B2_JNT_FREE) {
        memset(d->cdof_dot + 6*da, 0, 18 * sizeof(real));
        for (int a = 0; a < 3; a++) for (int c = 0; c < 6; c++) cv[c] += d->cdof[6*(da+a)+c] * d->qvel[da+a];
        for (int a = 3; a < 6; a++) cross_motion(d->cdof_dot + 6*(da+a), cv, d->cdof + 6*(da+a));
        for (int a = 3; a < 6; a++) for (int c = 0; c < 6; c++) cv[c] += d->cdof[6*(da+a)+c] * d->qvel[da+a];
      } else {
        cross_motion(d->cdof_dot + 6*da, cv, d->cdof + 6*da);
        for (int c = 0; c < 6; c++) cv[c] += d->cdof[6*da+c] * d->qvel[da];
      }
    }
    memcpy(d->cvel + 6*b, cv, sizeof(cv));
  }
}
static void passive(const RefModel *m, RefData *d) {
  for (int i = 0; i < NV; i++) d->qfrc_passive[i] = -MF(dof_damping)[i] * d->qvel[i];
  for (int j = 0; j < NJNT; j++) {
    real k = MF(jnt_stiffness)[j]; if (k == 0) continue;
    int t = MI(jnt_type)[j];
    if (t == B2_JNT_SLIDE || t == B2_JNT_HINGE) {
      int qa = MI(jnt_qposadr)[j];
      d->qfrc_passive[MI(jnt_dofadr)[j]] -= k * (d->qpos[qa] - MF(qpos_spring)[qa]);
    }
  }
}
/* mj_rne with flg_acc = 0: Coriolis + centrifugal + gravity */
static void rne_bias(const RefModel *m, RefData *d) {
  int nb = NBODY;
  real *cacc = d->cacc, *cfrc = d->cfrc;
  cacc[0] = cacc[1] = cacc[2] = 0; cacc[3] = -m->gravity[0]; cacc[4] = -m->gravity[1]; cacc[5] = -m->gravity[2];
  memset(cfrc, 0, 6 * sizeof(real));
  for (int b = 1; b < nb; b++) {
    real *a = cacc + 6*b; memcpy(a, cacc + 6*MI(body_parentid)[b], 6 * sizeof(real));
    int da = MI(body_dofadr)[b], dn = MI(body_dofnum)[b];
    for (int k = 0; k < dn; k++) for (int c = 0; c < 6; c++) a[c] += d->cdof_dot[6*(da+k)+c] * d->qvel[da+k];
    real t1[6], t2[6];
    mul_inert_vec(t1, d->cinert + 10*b, a);
    mul_inert_vec(t2, d->cinert + 10*b, d->cvel + 6*b);
    cross_force(cfrc + 6*b, d->cvel + 6*b, t2);
    for (int c = 0; c < 6; c++) cfrc[6*b+c] += t1[c];
  }
  for (int b = nb - 1; b > 0; b--) { int p = MI(body_parentid)[b]; if (p > 0) for (int c = 0; c < 6; c++) cfrc[6*p+c] += cfrc[6*b+c]; }
  for (int i = 0; i < NV; i++) {
    real s = 0; const real *f = cfrc + 6*MI(dof_bodyid)[i];
    for (int c = 0; c < 6; c++) s += d->cdof[6*i+c] * f[c];
    d->qfrc_bias[i] = s;
  }
}
static void reference_constraint(const RefModel *m, RefData *d) {
  int nv = NV;
  for (int i = 0; i < d->nefc; i++) {
    real v = 0; const real *J = d->efc_J + (size_t)i*nv;
    for (int k = 0; k < nv; k++) v += J[k]*d->qvel[k];
    d->efc_vel[i] = v;
    const real *kb = d->efc_KBIP + 4*i;
    d->efc_aref[i] = -kb[1]*v - kb[0]*kb[2]*(d->efc_pos[i] - d->efc_margin[i]);
  }
}
static void fwd_velocity(const RefModel *m, RefData *d) {
  com_vel(m, d); passive(m, d); reference_constraint(m, d); rne_bias(m, d);
}

/* ------------------------------------------------------------------ B.3 actuation / acceleration / solver */
static void fwd_actuation(const RefModel *m, RefData *d) {
  memset(d->qfrc_actuator, 0, sizeof(real) * NV);
  for (int i = 0; i < NU; i++) {
    real c = d->ctrl[i]; int dof = MI(act_dofid)[i];
    if (MI(act_ctrllimited)[i]) { real lo = MF(act_ctrlrange)[2*i], hi = MF(act_ctrlrange)[2*i+1]; c = c < lo ? lo : (c > hi ? hi : c); }
    /* joint transmission: length = gear*q, velocity = gear*qvel */
    int qa = MI(jnt_qposadr)[MI(dof_jntid)[dof]]; real gear = MF(act_gear)[i];
    real f = MF(act_gainprm)[i]*c + MF(act_biasprm)[3*i] + MF(act_biasprm)[3*i+1]*gear*d->qpos[qa] + MF(act_biasprm)[3*i+2]*gear*d->qvel[dof];
    if (MI(act_forcelimited)[i]) { real lo = MF(act_forcerange)[2*i], hi = MF(act_forcerange)[2*i+1]; f = f < lo ? lo : (f > hi ? hi : f); }
    d->actuator_force[i] = f;
    d->qfrc_actuator[dof] += gear * f;
  }
}
static void fwd_acceleration(const RefModel *m, RefData *d) {
  int nv = NV;
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i] + d->qfrc_actuator[i];
  /* Cartesian applied forces: qfrc += J' [torque-at-com-of-body]  (mj_xfrcAccumulate) */
  real *jp = d->scratch, *jr = jp + 3*nv;
  for (int b = 1; b < NBODY; b++) {
    const real *xf = d->xfrc_applied + 6*b;
    if (xf[0] == 0 && xf[1] == 0 && xf[2] == 0 && xf[3] == 0 && xf[4] == 0 && xf[5] == 0) continue;
    jac_point(m, d, jp, jr, d->xipos + 3*b, b);
    for (int i = 0; i < nv; i++) for (int a = 0; a < 3; a++) d->qfrc_smooth[i] += jp[a*nv+i]*xf[a] + jr[a*nv+i]*xf[3+a];
  }
  memcpy(d->qacc_smooth, d->qfrc_smooth, sizeof(real) * nv);
  chol_solve(d->L, d->qacc_smooth, nv);
}
/* efc_force from jar (mj_constraintUpdate, pyramidal/limit rows are one-sided quadratics) */
static real constraint_update(RefData *d, const real *jar) {
  real cost = 0;
  for (int i = 0; i < d->nefc; i++) {
    if (jar[i] < 0) { d->efc_force[i] = -d->efc_D[i]*jar[i]; cost += 0.5*d->efc_D[i]*jar[i]*jar[i]; }
    else d->efc_force[i] = 0;
  }
  return cost;
}
static void sol_pgs(const RefModel *m, RefData *d) {
  int ne = d->nefc; const real *AR = d->efc_AR;
  real scale = 1.0 / (m->meaninertia * (NV > 1 ? NV : 1));
  int it = 0;
  for (; it < m->d[B2D_iterations]; it++) {
    real improvement = 0;
    for (int i = 0; i < ne; i++) {
      real res = d->efc_b[i]; const real *row = AR + (size_t)i*ne;
      for (int j = 0; j < ne; j++) res += row[j]*d->efc_force[j];
      real old = d->efc_force[i], f = old - res / row[i];
      if (f < 0) f = 0;
      real delta = f - old, change = 0.5*delta*delta*row[i] + delta*res;
      if (change > 1e-10) { f = old; change = 0; }
      d->efc_force[i] = f;
      improvement -= change;
    }
    improvement *= scale;
    if (improvement < m->tolerance) { it++; break; }
  }
  d->solver_iter = it;
}
int ref_sol_newton(const RefModel *m, RefData *d); /* defined below */

/* mj_fwdConstraint (engine_forward.c, 3.x): every call ends by saving qacc into qacc_warmstart -- qacc_smooth when there are
 * no constraint rows -- so under RK4 stages 2-4 are warm-started from the previous *stage* and mj_forward moves the warm
 * start too.  warmstart_once_per_step = 1 restores the round-1 reading (saved once per mj_step, by the integrator). */
static void fwd_constraint_inner(const RefModel *m, RefData *d);
static void fwd_constraint(const RefModel *m, RefData *d) {
  fwd_constraint_inner(m, d);
  if (!d->warmstart_once_per_step) memcpy(d->qacc_warmstart, d->qacc, sizeof(real) * NV);
}
static void fwd_constraint_inner(const RefModel *m, RefData *d) {
  int nv = NV, ne = d->nefc;
  if (ne == 0) { memcpy(d->qacc, d->qacc_smooth, sizeof(real) * nv); memset(d->qfrc_constraint, 0, sizeof(real) * nv); d->solver_iter = 0; return; }
  for (int i = 0; i < ne; i++) {
    real s = 0; const real *J = d->efc_J + (size_t)i*nv;
    for (int k = 0; k < nv; k++) s += J[k]*d->qacc_smooth[k];
    d->efc_b[i] = s - d->efc_aref[i];
  }
  real *jar = (real *)malloc(sizeof(real) * ne);
  if (m->d[B2D_solver] == B2_SOLVER_PGS) {
    if (!d->disable_warmstart) {
      for (int i = 0; i < ne; i++) {
        real s = 0; const real *J = d->efc_J + (size_t)i*nv;
        for (int k = 0; k < nv; k++) s += J[k]*d->qacc_warmstart[k];
        jar[i] = s - d->efc_aref[i];
      }
      constraint_update(d, jar);
      real cost = 0;
      for (int i = 0; i < ne; i++) {
        real r = 0; const real *row = d->efc_AR + (size_t)i*ne;
        for (int j = 0; j < ne; j++) r += row[j]*d->efc_force[j];
        cost += 0.5*d->efc_force[i]*r + d->efc_force[i]*d->efc_b[i];
      }
      if (cost > 0) memset(d->efc_force, 0, sizeof(real) * ne);
    } else memset(d->efc_force, 0, sizeof(real) * ne);
    sol_pgs(m, d);
    for (int k = 0; k < nv; k++) {
      real s = 0; for (int i = 0; i < ne; i++) s += d->efc_J[(size_t)i*nv+k]*d->efc_force[i];
      d->qfrc_constraint[k] = s;
    }
    memcpy(d->qacc, d->qfrc_constraint, sizeof(real) * nv);
    chol_solve(d->L, d->qacc, nv);
    for (int k = 0; k < nv; k++) d->qacc[k] += d->qacc_smooth[k];
  } else {
    ref_sol_newton(m, d);
  }
  free(jar);
}

/* Newton solver (primal, mj_solNewton restated for one-sided quadratic rows):
 * minimise  1/2 (a-a_s)'M(a-a_s) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2
 * with exact Hessian H = M + J' diag(D active) J (dense Cholesky each iteration) and an exact
 * piecewise-quadratic line search.  Converged optimum is what Newton tasks are compared on. */
int ref_sol_newton(const RefModel *m, RefData *d) {
  int nv = NV, ne = d->nefc;
  real *a = d->qacc, *Ma = (real *)malloc(sizeof(real) * nv), *grad = (real *)malloc(sizeof(real) * nv);
  real *H = (real *)malloc(sizeof(real) * nv * nv), *LH = (real *)malloc(sizeof(real) * nv * nv);
  real *jar = (real *)malloc(sizeof(real) * ne), *jv = (real *)malloc(sizeof(real) * ne), *search = (real *)malloc(sizeof(real) * nv);
  real *Mv = (real *)malloc(sizeof(real) * nv);
  real scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  /* warm start: better of qacc_warmstart and qacc_smooth */
  real cost_ws = 0, cost_sm = 0;
  for (int pass = 0; pass < 2; pass++) {
    const real *x = pass == 0 ? d->qacc_warmstart : d->qacc_smooth; real c = 0;
    for (int i = 0; i < ne; i++) {
      real s = 0; const real *J = d->efc_J + (size_t)i*nv; for (int k = 0; k < nv; k++) s += J[k]*x[k];
      s -= d->efc_aref[i]; if (s < 0) c += 0.5*d->efc_D[i]*s*s;
    }
    for (int i = 0; i < nv; i++) {
      real mx = 0; for (int k = 0; k < nv; k++) mx += d->M[i*nv+k]*x[k];
      c += 0.5*(mx - d->qfrc_smooth[i])*(x[i] - d->qacc_smooth[i]);
    }
    if (pass == 0) cost_ws = c; else cost_sm = c;
  }
  memcpy(a, (!d->disable_warmstart && cost_ws < cost_sm) ? d->qacc_warmstart : d->qacc_smooth, sizeof(real) * nv);
  int it = 0;
  for (; it < m->d[B2D_iterations]; it++) {
    for (int i = 0; i < ne; i++) { real s = 0; const real *J = d->efc_J + (size_t)i*nv; for (int k = 0; k < nv; k++) s += J[k]*a[k]; jar[i] = s - d->efc_aref[i]; }
    for (int i = 0; i < nv; i++) { real s = 0; for (int k = 0; k < nv; k++) s += d->M[i*nv+k]*a[k]; Ma[i] = s; grad[i] = s - d->qfrc_smooth[i]; }
    memcpy(H, d->M, sizeof(real) * nv * nv);
    for (int i = 0; i < ne; i++) if (jar[i] < 0) {
      const real *J = d->efc_J + (size_t)i*nv; real D = d->efc_D[i];
      for (int r = 0; r < nv; r++) { if (J[r] == 0) continue; grad[r] += D*jar[i]*J[r]; for (int c = 0; c < nv; c++) H[r*nv+c] += D*J[r]*J[c]; }
    }
    real gn = 0; for (int i = 0; i < nv; i++) gn += grad[i]*grad[i];
    if (scale * sqrt(gn) < m->tolerance) break;
    cholesky(LH, H, nv);
    for (int i = 0; i < nv; i++) search[i] = -grad[i];
    chol_solve(LH, search, nv);
    for (int i = 0; i < ne; i++) { real s = 0; const real *J = d->efc_J + (size_t)i*nv; for (int k = 0; k < nv; k++) s += J[k]*search[k]; jv[i] = s; }
    for (int i = 0; i < nv; i++) { real s = 0; for (int k = 0; k < nv; k++) s += d->M[i*nv+k]*search[k]; Mv[i] = s; }
    /* exact 1-D minimisation of the piecewise quadratic by safeguarded Newton on its derivative */
    real q1 = 0, q2 = 0;
    for (int i = 0; i < nv; i++) { q1 += search[i]*(Ma[i] - d->qfrc_smooth[i]); q2 += search[i]*Mv[i]; }
    real alpha = 0, lo = 0, hi = -1;
    for (int ls = 0; ls < 100; ls++) {
      real d1 = q1 + alpha*q2, d2 = q2;
      for (int i = 0; i < ne; i++) { real x = jar[i] + alpha*jv[i]; if (x < 0) { d1 += d->efc_D[i]*x*jv[i]; d2 += d->efc_D[i]*jv[i]*jv[i]; } }
      if (fabs(d1) < 1e-14 * (1 + fabs(q1))) break;
      if (d1 < 0) lo = alpha; else hi = alpha;
      real na = alpha - d1 / d2;
      if (hi > 0 && (na <= lo || na >= hi)) na = 0.5*(lo + hi);
      if (na < 0) na = 0;
      if (fabs(na - alpha) < 1e-15 * (1 + fabs(alpha))) { alpha = na; break; }
      alpha = na;
    }
    real impr = 0;
    for (int i = 0; i < nv; i++) { a[i] += alpha*search[i]; impr += fabs(alpha*search[i]); }
    if (impr == 0) { it++; break; }
  }
  d->solver_iter = it;
  for (int i = 0; i < ne; i++) { real s = 0; const real *J = d->efc_J + (size_t)i*nv; for (int k = 0; k < nv; k++) s += J[k]*a[k]; s -= d->efc_aref[i]; d->efc_force[i] = s < 0 ? -d->efc_D[i]*s : 0; }
  for (int k = 0; k < nv; k++) { real s = 0; for (int i = 0; i < ne; i++) s += d->efc_J[(size_t)i*nv+k]*d->efc_force[i]; d->qfrc_constraint[k] = s; }
  free(Ma); free(grad); free(H); free(LH); free(jar); free(jv); free(search); free(Mv);
  return it;
}

void ref_forward(const RefModel *m, RefData *d) {
  fwd_position(m, d); fwd_velocity(m, d); fwd_actuation(m, d); fwd_acceleration(m, d); fwd_constraint(m, d);
}

/* ------------------------------------------------------------------ B.7 integrators */
static void integrate_pos(const RefModel *m, real *qpos, const real *qvel, real h) {
  for (int j = 0; j < NJNT; j++) {
    int qa = MI(jnt_qposadr)[j], da = MI(jnt_dofadr)[j];
    if (MI(jnt_type)[j] == B2_JNT_FREE) {
      for (int k = 0; k < 3; k++) qpos[qa+k] += h * qvel[da+k];
      quat_integrate(qpos + qa + 3, qvel + da + 3, h);
    } else qpos[qa] += h * qvel[da];
  }
}
static void advance(const RefModel *m, RefData *d, const real *qacc, const real *qvel_for_pos) {
  real h = m->timestep;
  for (int i = 0; i < NV; i++) d->qvel[i] += h * qacc[i];
  integrate_pos(m, d->qpos, qvel_for_pos ? qvel_for_pos : d->qvel, h);
  d->time += h;
  memcpy(d->qacc_warmstart, d->qacc, sizeof(real) * NV);
}
static void euler(const RefModel *m, RefData *d) {
  int nv = NV; real h = m->timestep;
  int damped = 0; for (int i = 0; i < nv; i++) if (MF(dof_damping)[i] > 0) damped = 1;
  if (!damped || d->disable_eulerdamp) { advance(m, d, d->qacc, NULL); return; }
  real *H = d->scratch, *LH = H + nv*nv, *q = LH + nv*nv;
  memcpy(H, d->M, sizeof(real) * nv * nv);
  for (int i = 0; i < nv; i++) { H[i*nv+i] += h * MF(dof_damping)[i]; q[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
  cholesky(LH, H, nv); chol_solve(LH, q, nv);
  advance(m, d, q, NULL);
}
static void rk4(const RefModel *m, RefData *d) {
  int nq = NQ, nv = NV; real h = m->timestep, t0 = d->time;
  static const real A[3][3] = {{0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 1}}, Bw[4] = {1.0/6, 1.0/3, 1.0/3, 1.0/6}, C[3] = {0.5, 0.5, 1.0};
  real *q0 = (real *)malloc(sizeof(real) * (nq + nv)), *v0 = q0 + nq;
  real *Fv = (real *)malloc(sizeof(real) * 8 * nv), *Fa = Fv + 4*nv, *dv = (real *)malloc(sizeof(real) * 2 * nv), *da = dv + nv;
  memcpy(q0, d->qpos, sizeof(real) * nq); memcpy(v0, d->qvel, sizeof(real) * nv);
  memcpy(Fv, d->qvel, sizeof(real) * nv); memcpy(Fa, d->qacc, sizeof(real) * nv);
  for (int i = 1; i < 4; i++) {
    for (int k = 0; k < nv; k++) { dv[k] = 0; da[k] = 0; for (int j = 0; j < i; j++) { dv[k] += A[i-1][j]*Fv[j*nv+k]; da[k] += A[i-1][j]*Fa[j*nv+k]; } }
    memcpy(d->qpos, q0, sizeof(real) * nq); integrate_pos(m, d->qpos, dv, h);
    for (int k = 0; k < nv; k++) d->qvel[k] = v0[k] + h*da[k];
    d->time = t0 + h*C[i-1];
    ref_forward(m, d);
    memcpy(Fv + i*nv, d->qvel, sizeof(real) * nv); memcpy(Fa + i*nv, d->qacc, sizeof(real) * nv);
  }
  for (int k = 0; k < nv; k++) { dv[k] = 0; da[k] = 0; for (int j = 0; j < 4; j++) { dv[k] += Bw[j]*Fv[j*nv+k]; da[k] += Bw[j]*Fa[j*nv+k]; } }
  memcpy(d->qpos, q0, sizeof(real) * nq); memcpy(d->qvel, v0, sizeof(real) * nv); d->time = t0;
  advance(m, d, da, dv);
  free(q0); free(Fv); free(dv);
}

static int bad(const real *x, int n) { for (int i = 0; i < n; i++) if (!(x[i] == x[i]) || x[i] > MAXVAL || x[i] < -MAXVAL) return 1; return 0; }

/* mj_step: checkPos/Vel -> forward -> checkAcc -> integrate */
void ref_step(const RefModel *m, RefData *d) {
  if (bad(d->qpos, NQ) || bad(d->qvel, NV)) { ref_reset_data(m, d); d->nwarn_bad++; }
  ref_forward(m, d);
  if (bad(d->qacc, NV)) { ref_reset_data(m, d); d->nwarn_bad++; ref_forward(m, d); }
  if (m->d[B2D_integrator] == B2_INT_RK4) rk4(m, d); else euler(m, d);
}
void ref_step_n(const RefModel *m, RefData *d, int n) { for (int i = 0; i < n; i++) ref_step(m, d); }

/* ------------------------------------------------------------------ accessors for ctypes */
#define FIELD(nm, ptr, cnt) if (!strcmp(name, nm)) { *n = (cnt); return (ptr); }
double *ref_field(const RefModel *m, RefData *d, const char *name, int *n) {
  FIELD("qpos", d->qpos, NQ) FIELD("qvel", d->qvel, NV) FIELD("ctrl", d->ctrl, NU)
  FIELD("qfrc_applied", d->qfrc_applied, NV) FIELD("xfrc_applied", d->xfrc_applied, 6*NBODY)
  FIELD("qacc", d->qacc, NV) FIELD("qacc_warmstart", d->qacc_warmstart, NV)
  FIELD("xpos", d->xpos, 3*NBODY) FIELD("xquat", d->xquat, 4*NBODY) FIELD("xmat", d->xmat, 9*NBODY)
  FIELD("xipos", d->xipos, 3*NBODY) FIELD("ximat", d->ximat, 9*NBODY)
  FIELD("geom_xpos", d->geom_xpos, 3*NGEOM) FIELD("geom_xmat", d->geom_xmat, 9*NGEOM)
  FIELD("site_xpos", d->site_xpos, 3*NSITE) FIELD("subtree_com", d->subtree_com, 3*NBODY)
  FIELD("cvel", d->cvel, 6*NBODY) FIELD("cdof", d->cdof, 6*NV) FIELD("cinert", d->cinert, 10*NBODY)
  FIELD("M", d->M, NV*NV) FIELD("qfrc_bias", d->qfrc_bias, NV) FIELD("qfrc_passive", d->qfrc_passive, NV)
  FIELD("qfrc_actuator", d->qfrc_actuator, NV) FIELD("qfrc_smooth", d->qfrc_smooth, NV)
  FIELD("qacc_smooth", d->qacc_smooth, NV) FIELD("qfrc_constraint", d->qfrc_constraint, NV)
  FIELD("efc_force", d->efc_force, d->nefc) FIELD("efc_pos", d->efc_pos, d->nefc) FIELD("efc_R", d->efc_R, d->nefc)
  FIELD("efc_aref", d->efc_aref, d->nefc) FIELD("efc_b", d->efc_b, d->nefc) FIELD("efc_D", d->efc_D, d->nefc)
  FIELD("efc_J", d->efc_J, d->nefc*NV) FIELD("efc_AR", d->efc_AR, d->nefc*d->nefc)
  FIELD("time", &d->time, 1)
  *n = 0; return NULL;
}
int ref_ncon(const RefData *d) { return d->ncon; }
int ref_nefc(const RefData *d) { return d->nefc; }
int ref_solver_iter(const RefData *d) { return d->solver_iter; }
int ref_nwarn(const RefData *d) { return d->nwarn_bad; }
void ref_set_warmstart_mode(RefData *d, int once_per_step) { d->warmstart_once_per_step = once_per_step; }
void ref_set_flags(RefData *d, int disable_eulerdamp, int disable_warmstart) { d->disable_eulerdamp = disable_eulerdamp; d->disable_warmstart = disable_warmstart; }
/* contact k -> (geom1, geom2, dist, pos[3], frame[9], friction[5]) */
void ref_contact(const RefData *d, int k, int *geoms, double *out) {
  const RefContact *c = d->contact + k;
  geoms[0] = c->geom1; geoms[1] = c->geom2; geoms[2] = c->dim; geoms[3] = c->efc_address;
  out[0] = c->dist; memcpy(out + 1, c->pos, 3 * sizeof(real)); memcpy(out + 4, c->frame, 9 * sizeof(real));
  memcpy(out + 13, c->friction, 5 * sizeof(real));
}
