/*
 * oracle/bullet_restatement.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * fp64, single-environment CPU restatement of the arithmetic the reference (ascentai/diy-gym) delegates to its
 * third-party physics dependency `pybullet` (unpinned: /root/reference/requirements.txt:3) on the hot path
 * `DIYGym.step` (diy_gym/diy_gym.py:187-209): `p.stepSimulation()` (:207) with numSubSteps=2, fixedTimeStep=1/240,
 * numSolverIterations=150 (:76-79), plus the add-on arithmetic of diy_gym/addons/ and the camera of
 * diy_gym/addons/sensors/camera.py:58-92.
 *
 * PARITY UNPINNED: pybullet/Bullet are absent from this container and the reference's own tests pin no number
 * on this path (diy_gym/tests/test_environment.py:34,40).  The physics below restates the published algorithms
 * (Featherstone articulated-body algorithm, projected Gauss-Seidel on multibody constraint rows, damped least
 * squares IK) with the engine semantics listed in SURVEY.md Appendix A, every one of which is [RECALLED-UNVERIFIED].
 * It is pinned only against analytic known-answer tests (tests/test_oracle_kat.py).  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use this file.
 *
 * Formulation (own restatement): link frames are inertial (COM) frames; spatial vectors are [angular; linear]
 * at the COM origin in link coordinates; generalized velocity of a body is [omega_world(3), v_com_world(3), qd].
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "scene_sections.h"

#define MAXROWDOF 64

static double g_flops = 0.0;
#define FL(n) (g_flops += (n))

/* ---------------------------------------------------------------- small math ---------------------------- */
static inline void v_set(double* o, double x, double y, double z) { o[0] = x; o[1] = y; o[2] = z; }
static inline void v_cpy(double* o, const double* a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
static inline void v_add(double* o, const double* a, const double* b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; FL(3); }
static inline void v_sub(double* o, const double* a, const double* b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; FL(3); }
static inline void v_scale(double* o, const double* a, double s) { o[0] = a[0] * s; o[1] = a[1] * s; o[2] = a[2] * s; FL(3); }
static inline void v_madd(double* o, const double* a, double s) { o[0] += a[0] * s; o[1] += a[1] * s; o[2] += a[2] * s; FL(6); }
static inline double v_dot(const double* a, const double* b) { FL(5); return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double v_len(const double* a) { return sqrt(v_dot(a, a)); }
static inline void v_cross(double* o, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z; FL(9);
}
/* mat3 row-major */
static inline void m_vec(double* o, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z; FL(15);
}
static inline void mT_vec(double* o, const double* m, const double* v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z; FL(15);
}
static inline void m_mul(double* o, const double* a, const double* b) {
  double t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(o, t, sizeof t); FL(45);
}
static inline void m_T(double* o, const double* a) {
  double t[9] = {a[0], a[3], a[6], a[1], a[4], a[7], a[2], a[5], a[8]};
  memcpy(o, t, sizeof t);
}
static inline void m_mulT(double* o, const double* a, const double* b) { double bt[9]; m_T(bt, b); m_mul(o, a, bt); }   /* a b^T */
static inline void mT_mul(double* o, const double* a, const double* b) { double at[9]; m_T(at, a); m_mul(o, at, b); }   /* a^T b */
static inline void m_skew(double* o, const double* r) { o[0] = 0; o[1] = -r[2]; o[2] = r[1]; o[3] = r[2]; o[4] = 0; o[5] = -r[0]; o[6] = -r[1]; o[7] = r[0]; o[8] = 0; }
/* quaternions xyzw */
static inline void q_to_mat(double* m, const double* q) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double n = 1.0 / sqrt(x * x + y * y + z * z + w * w); x *= n; y *= n; z *= n; w *= n;
  m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - z * w); m[2] = 2 * (x * z + y * w);
  m[3] = 2 * (x * y + z * w); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - x * w);
  m[6] = 2 * (x * z - y * w); m[7] = 2 * (y * z + x * w); m[8] = 1 - 2 * (x * x + y * y); FL(40);
}
static inline void q_mul(double* o, const double* a, const double* b) {
  double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w; FL(28);
}
static inline void q_norm(double* q) { double n = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n; }
static void mat_to_q(double* q, const double* m) {
  double t = m[0] + m[4] + m[8];
  if (t > 0) { double s = sqrt(t + 1.0) * 2; q[3] = 0.25 * s; q[0] = (m[7] - m[5]) / s; q[1] = (m[2] - m[6]) / s; q[2] = (m[3] - m[1]) / s; }
  else if (m[0] > m[4] && m[0] > m[8]) { double s = sqrt(1.0 + m[0] - m[4] - m[8]) * 2; q[3] = (m[7] - m[5]) / s; q[0] = 0.25 * s; q[1] = (m[1] + m[3]) / s; q[2] = (m[2] + m[6]) / s; }
  else if (m[4] > m[8]) { double s = sqrt(1.0 + m[4] - m[0] - m[8]) * 2; q[3] = (m[2] - m[6]) / s; q[0] = (m[1] + m[3]) / s; q[1] = 0.25 * s; q[2] = (m[5] + m[7]) / s; }
  else { double s = sqrt(1.0 + m[8] - m[0] - m[4]) * 2; q[3] = (m[3] - m[1]) / s; q[0] = (m[2] + m[6]) / s; q[1] = (m[5] + m[7]) / s; q[2] = 0.25 * s; }
  q_norm(q);
}
static void q_from_axis_angle(double* q, const double* a, double ang) { double s = sin(0.5 * ang); q[0] = a[0] * s; q[1] = a[1] * s; q[2] = a[2] * s; q[3] = cos(0.5 * ang); FL(6); }
/* R = Rz(yaw) Ry(pitch) Rx(roll): pybullet getQuaternionFromEuler (used at diy_gym/model.py:54, respawn.py:39) */
static void q_from_euler(double* q, const double* rpy) {
  double cr = cos(rpy[0] / 2), sr = sin(rpy[0] / 2), cp = cos(rpy[1] / 2), sp = sin(rpy[1] / 2), cy = cos(rpy[2] / 2), sy = sin(rpy[2] / 2);
  q[0] = sr * cp * cy - cr * sp * sy; q[1] = cr * sp * cy + sr * cp * sy; q[2] = cr * cp * sy - sr * sp * cy; q[3] = cr * cp * cy + sr * sp * sy;
}
/* pybullet getEulerFromQuaternion (used at diy_gym/addons/sensors/object_state_sensor.py:70); SURVEY App. A.5 */
static void euler_from_q(double* rpy, const double* q) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double sarg = -2 * (x * z - w * y);
  if (sarg <= -0.99999) { rpy[0] = 0; rpy[1] = -0.5 * M_PI; rpy[2] = 2 * atan2(x, -y); }
  else if (sarg >= 0.99999) { rpy[0] = 0; rpy[1] = 0.5 * M_PI; rpy[2] = 2 * atan2(-x, y); }
  else {
    rpy[0] = atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z);
    rpy[1] = asin(sarg);
    rpy[2] = atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
  }
}

/* ---------------------------------------------------------------- world --------------------------------- */
typedef struct { int fa, fb; double pa[3], pb[3], n[3], dist, mu; } Contact;
typedef struct {
  int bodyA, bodyB;            /* -1 = none */
  double JA[MAXROWDOF], JB[MAXROWDOF], MA[MAXROWDOF], MB[MAXROWDOF];
  double rhs, diag_inv, lo, hi, applied, mu;
  int parent;                  /* friction rows: index of the normal row, else -1 */
  int motor_dof;               /* motor rows: global dof whose applied impulse is recorded, else -1 */
  int unit_dof;                /* rows whose Jacobian is +/- one joint coordinate: local dof index, else -1 */
} Row;

typedef struct DgoWorld {
  int nb, nl, nd, ns, nv, npair, ncam, nop, nframes, S, P, substeps, iters, maxc, hot_start, ik_iters, need_react, ncons, sem;
  const int32_t* cons_i; const double* cons_f;
  const double* hull_f;   /* reduced convex hulls of mesh collision shapes (shape_i[4..7]) */
  int32_t* ibuf; double* fbuf;
  const int32_t *hi, *body_i, *link_i, *shape_i, *pair_i, *vis_i, *op_i, *oparg_i, *cam_i;
  const double *hf, *body_f, *link_f, *shape_f, *vis_f, *oparg_f, *param_def, *state_def, *cam_f;
  double *state, *param;
  double *Rw, *pw;             /* per frame: world_from_frame rotation, COM position */
  double *E, *r, *Sa, *Sl;     /* per link: child_from_parent rotation, child origin in parent coords, motion subspace */
  double *w, *v;               /* per frame: spatial velocity in frame coords */
  double *c, *pA, *IAa, *IAb, *IAc, *U, *D, *u, *acc, *I0inv, *qdd;
  Contact* contacts; int ncontacts;
  Row* rows; int nrows;
  double* dv;                  /* per body, MAXROWDOF each */
  uint32_t seed; int env_id;
  long dropped;   /* contacts lost to the max_contacts cap since creation */
} DgoWorld;

#define HI(W, k) ((W)->hi[HI_##k])
#define HF(W, k) ((W)->hf[HF_##k])
#define ST(W, k) ((W)->state + HI(W, k))
#define PR(W, k) ((W)->param + HI(W, k))

static const int32_t* sec_i(const int32_t* ib, int s) { return ib + ib[2 + 3 * s + 1]; }
static const double* sec_f(const int32_t* ib, const double* fb, int s) { return fb + ib[2 + 3 * s + 1]; }

DgoWorld* dgo_create(const int32_t* ibuf, int ni, const double* fbuf, int nf) {
  if (ni < 2 || ibuf[0] != DG_SCENE_MAGIC || ibuf[1] != DG_NSECTIONS) return NULL;
  DgoWorld* W = (DgoWorld*)calloc(1, sizeof(DgoWorld));
  W->ibuf = (int32_t*)malloc(sizeof(int32_t) * (size_t)ni); memcpy(W->ibuf, ibuf, sizeof(int32_t) * (size_t)ni);
  W->fbuf = (double*)malloc(sizeof(double) * (size_t)(nf > 0 ? nf : 1)); memcpy(W->fbuf, fbuf, sizeof(double) * (size_t)nf);
  const int32_t* ib = W->ibuf; const double* fb = W->fbuf;
  W->hi = sec_i(ib, SEC_HDR_I); W->hf = sec_f(ib, fb, SEC_HDR_F);
  W->body_i = sec_i(ib, SEC_BODY_I); W->body_f = sec_f(ib, fb, SEC_BODY_F);
  W->link_i = sec_i(ib, SEC_LINK_I); W->link_f = sec_f(ib, fb, SEC_LINK_F);
  W->shape_i = sec_i(ib, SEC_SHAPE_I); W->shape_f = sec_f(ib, fb, SEC_SHAPE_F);
  W->pair_i = sec_i(ib, SEC_PAIR_I); W->vis_i = sec_i(ib, SEC_VIS_I); W->vis_f = sec_f(ib, fb, SEC_VIS_F);
  W->op_i = sec_i(ib, SEC_OP_I); W->oparg_i = sec_i(ib, SEC_OPARG_I); W->oparg_f = sec_f(ib, fb, SEC_OPARG_F);
  W->param_def = sec_f(ib, fb, SEC_PARAM_DEFAULT); W->state_def = sec_f(ib, fb, SEC_STATE_DEFAULT);
  W->cam_i = sec_i(ib, SEC_CAM_I); W->cam_f = sec_f(ib, fb, SEC_CAM_F);
  W->nb = HI(W, nb); W->nl = HI(W, nl); W->nd = HI(W, nd); W->ns = HI(W, ns); W->nv = HI(W, nv); W->npair = HI(W, npair);
  W->ncam = HI(W, ncam); W->nop = HI(W, nop); W->nframes = HI(W, nframes); W->S = HI(W, S); W->P = HI(W, P);
  W->substeps = HI(W, substeps); W->iters = HI(W, iterations); W->maxc = HI(W, max_contacts); W->hot_start = HI(W, hot_start);
  W->ik_iters = HI(W, ik_iters); W->sem = HI(W, semantics);   /* SEM_* switches (compiler/scene.py SEMANTICS) */
  W->ncons = HI(W, ncons); W->cons_i = sec_i(ib, SEC_CONS_I); W->cons_f = sec_f(ib, fb, SEC_CONS_F); W->hull_f = sec_f(ib, fb, SEC_HULL_F);
  W->need_react = HI(W, S_STEP) > HI(W, S_JREACT);   /* the state row holds reaction wrenches only when a sensor asked for them */
  int nf_ = W->nframes, nl = W->nl > 0 ? W->nl : 1;
#define ALLOC(p, n) W->p = (double*)calloc((size_t)(n) > 0 ? (size_t)(n) : 1, sizeof(double))
  ALLOC(state, W->S); ALLOC(param, W->P); ALLOC(Rw, 9 * nf_); ALLOC(pw, 3 * nf_); ALLOC(E, 9 * nl); ALLOC(r, 3 * nl);
  ALLOC(Sa, 3 * nl); ALLOC(Sl, 3 * nl); ALLOC(w, 3 * nf_); ALLOC(v, 3 * nf_); ALLOC(c, 6 * nl); ALLOC(pA, 6 * nf_);
  ALLOC(IAa, 9 * nf_); ALLOC(IAb, 9 * nf_); ALLOC(IAc, 9 * nf_); ALLOC(U, 6 * nl); ALLOC(D, nl); ALLOC(u, nl); ALLOC(acc, 6 * nf_);
  ALLOC(I0inv, 36 * W->nb); ALLOC(qdd, W->nd); ALLOC(dv, MAXROWDOF * W->nb);
  W->contacts = (Contact*)calloc((size_t)W->maxc + 1, sizeof(Contact));
  W->rows = (Row*)calloc((size_t)(3 * W->nd + 3 * W->maxc + 1), sizeof(Row));
  memcpy(W->state, W->state_def, sizeof(double) * (size_t)W->S);
  memcpy(W->param, W->param_def, sizeof(double) * (size_t)W->P);
  W->seed = 1234u; W->env_id = 0;
  return W;
}
void dgo_destroy(DgoWorld* W) {
  if (!W) return;
  double* ptrs[] = {W->state, W->param, W->Rw, W->pw, W->E, W->r, W->Sa, W->Sl, W->w, W->v, W->c, W->pA, W->IAa, W->IAb, W->IAc,
                    W->U, W->D, W->u, W->acc, W->I0inv, W->qdd, W->dv, W->fbuf};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
  free(W->contacts); free(W->rows); free(W->ibuf); free(W);
}
double* dgo_state(DgoWorld* W) { return W->state; }
double* dgo_params(DgoWorld* W) { return W->param; }
int dgo_state_size(DgoWorld* W) { return W->S; }
int dgo_param_size(DgoWorld* W) { return W->P; }
void dgo_set_seed(DgoWorld* W, uint32_t seed, int env_id) { W->seed = seed; W->env_id = env_id; }
double dgo_flops(int reset) { double f = g_flops; if (reset) g_flops = 0; return f; }
int dgo_num_contacts(DgoWorld* W) { return W->ncontacts; }
long dgo_contacts_dropped(DgoWorld* W) { return W->dropped; }
void dgo_get_contact(DgoWorld* W, int i, double* out) {  /* fa, fb, pa3, pb3, n3, dist, mu */
  const Contact* c = &W->contacts[i]; out[0] = c->fa; out[1] = c->fb; memcpy(out + 2, c->pa, 24); memcpy(out + 5, c->pb, 24); memcpy(out + 8, c->n, 24); out[11] = c->dist; out[12] = c->mu;
}

/* ---------------------------------------------------------------- kinematics ---------------------------- */
static inline int frame_of_link(const DgoWorld* W, int gl) { return W->nb + gl; }
static inline int parent_frame(const DgoWorld* W, int gl) {
  const int32_t* li = W->link_i + DG_LINK_I_W * gl;
  return li[1] < 0 ? li[0] : W->nb + li[1];
}
/* joint transform of link gl at coordinate q: E = child_from_parent rotation, r = child COM in parent coords */
static void joint_xform(const DgoWorld* W, int gl, double q, double* E, double* r) {
  const int32_t* li = W->link_i + DG_LINK_I_W * gl; const double* lf = W->link_f + DG_LINK_F_W * gl;
  double R0[9], Rrel[9], tmp[3];
  q_to_mat(R0, lf);
  if (li[2] == 1) { double qa[4], Ra[9]; q_from_axis_angle(qa, lf + 10, q); q_to_mat(Ra, qa); m_mul(Rrel, R0, Ra); m_vec(tmp, Rrel, lf + 7); }
  else if (li[2] == 2) { memcpy(Rrel, R0, sizeof R0); double dd[3] = {lf[7] + lf[10] * q, lf[8] + lf[11] * q, lf[9] + lf[12] * q}; m_vec(tmp, R0, dd); }
  else { memcpy(Rrel, R0, sizeof R0); m_vec(tmp, R0, lf + 7); }
  v_add(r, lf + 4, tmp);
  m_T(E, Rrel);
}
/* world poses of every frame from (base pose, q); also E, r, S per link */
void dgo_forward_kinematics(DgoWorld* W) {
  const double *bpos = ST(W, S_BPOS), *bquat = ST(W, S_BQUAT), *q = ST(W, S_Q);
  for (int b = 0; b < W->nb; b++) { q_to_mat(W->Rw + 9 * b, bquat + 4 * b); v_cpy(W->pw + 3 * b, bpos + 3 * b); }
  for (int gl = 0; gl < W->nl; gl++) {
    const int32_t* li = W->link_i + DG_LINK_I_W * gl; const double* lf = W->link_f + DG_LINK_F_W * gl;
    int f = frame_of_link(W, gl), pf = parent_frame(W, gl);
    double qq = li[3] >= 0 ? q[li[3]] : 0.0;
    joint_xform(W, gl, qq, W->E + 9 * gl, W->r + 3 * gl);
    m_mulT(W->Rw + 9 * f, W->Rw + 9 * pf, W->E + 9 * gl);
    double t[3]; m_vec(t, W->Rw + 9 * pf, W->r + 3 * gl); v_add(W->pw + 3 * f, W->pw + 3 * pf, t);
    if (li[2] == 1) { v_cpy(W->Sa + 3 * gl, lf + 10); v_cross(W->Sl + 3 * gl, lf + 10, lf + 7); }
    else if (li[2] == 2) { v_set(W->Sa + 3 * gl, 0, 0, 0); v_cpy(W->Sl + 3 * gl, lf + 10); }
    else { v_set(W->Sa + 3 * gl, 0, 0, 0); v_set(W->Sl + 3 * gl, 0, 0, 0); }
  }
}
/* spatial velocities of every frame in frame coordinates */
static void velocities(DgoWorld* W) {
  const double *bvel = ST(W, S_BVEL), *bom = ST(W, S_BOMEGA), *qd = ST(W, S_QD);
  for (int b = 0; b < W->nb; b++) { mT_vec(W->w + 3 * b, W->Rw + 9 * b, bom + 3 * b); mT_vec(W->v + 3 * b, W->Rw + 9 * b, bvel + 3 * b); }
  for (int gl = 0; gl < W->nl; gl++) {
    const int32_t* li = W->link_i + DG_LINK_I_W * gl;
    int f = frame_of_link(W, gl), pf = parent_frame(W, gl);
    double t[3], t2[3];
    m_vec(W->w + 3 * f, W->E + 9 * gl, W->w + 3 * pf);
    v_cross(t, W->w + 3 * pf, W->r + 3 * gl); v_add(t2, W->v + 3 * pf, t); m_vec(W->v + 3 * f, W->E + 9 * gl, t2);
    if (li[3] >= 0) { double qdi = qd[li[3]]; v_madd(W->w + 3 * f, W->Sa + 3 * gl, qdi); v_madd(W->v + 3 * f, W->Sl + 3 * gl, qdi); }
  }
}
/* write the link pose / velocity cache that sensors read (what getLinkState reports after a step) */
static void write_link_cache(DgoWorld* W) {
  dgo_forward_kinematics(W); velocities(W);
  double *lpos = ST(W, S_LPOS), *lquat = ST(W, S_LQUAT), *lvel = ST(W, S_LVEL), *lom = ST(W, S_LOMEGA);
  for (int gl = 0; gl < W->nl; gl++) {
    int f = frame_of_link(W, gl);
    v_cpy(lpos + 3 * gl, W->pw + 3 * f); mat_to_q(lquat + 4 * gl, W->Rw + 9 * f);
    m_vec(lvel + 3 * gl, W->Rw + 9 * f, W->v + 3 * f); m_vec(lom + 3 * gl, W->Rw + 9 * f, W->w + 3 * f);
  }
}
void dgo_refresh(DgoWorld* W) { write_link_cache(W); }

/* ---------------------------------------------------------------- articulated-body algorithm ------------ */
/* I (3 blocks A,B,C) times motion vector (a;l) -> force (n;f):  n = A a + B l, f = B^T a + C l */
static void ia_mul(const double* A, const double* B, const double* C, const double* a, const double* l, double* n, double* f) {
  double t[3]; m_vec(n, A, a); m_vec(t, B, l); v_add(n, n, t); mT_vec(f, B, a); m_vec(t, C, l); v_add(f, f, t);
}
static int solve_sym(double* M, int n, double* inv) {  /* Gauss-Jordan with partial pivoting, M destroyed */
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) inv[i * n + j] = (i == j);
  for (int c = 0; c < n; c++) {
    int p = c; for (int r2 = c + 1; r2 < n; r2++) if (fabs(M[r2 * n + c]) > fabs(M[p * n + c])) p = r2;
    if (fabs(M[p * n + c]) < 1e-300) return -1;
    if (p != c) for (int j = 0; j < n; j++) { double t = M[c * n + j]; M[c * n + j] = M[p * n + j]; M[p * n + j] = t; t = inv[c * n + j]; inv[c * n + j] = inv[p * n + j]; inv[p * n + j] = t; }
    double d = 1.0 / M[c * n + c];
    for (int j = 0; j < n; j++) { M[c * n + j] *= d; inv[c * n + j] *= d; }
    for (int r2 = 0; r2 < n; r2++) if (r2 != c) { double fct = M[r2 * n + c]; if (fct != 0) for (int j = 0; j < n; j++) { M[r2 * n + j] -= fct * M[c * n + j]; inv[r2 * n + j] -= fct * inv[c * n + j]; } }
    FL(4 * n * n);
  }
  return 0;
}
/* forward dynamics of body b: fills qdd and acc (spatial acceleration per frame), caches U, D, I0inv */
static void aba_body(DgoWorld* W, int b, const double* tau_damp) {
  const int32_t* bi = W->body_i + DG_BODY_I_W * b;
  int kind = bi[0], l0 = bi[1], nlb = bi[2];
  if (kind == 0) return;
  const double *mass = PR(W, P_MASS), *inertia = PR(W, P_INERTIA);
  const double *extf = ST(W, S_EXTF), *extt = ST(W, S_EXTT), *jtq = ST(W, S_JTORQUE), *qd = ST(W, S_QD);
  double kl = PR(W, P_LINDAMP)[b], ka = PR(W, P_ANGDAMP)[b];
  double grav[3] = {HF(W, gx), HF(W, gy), HF(W, gz)};
  /* pass 1: bias forces and rigid-body inertias */
  for (int k = -1; k < nlb; k++) {
    int f = k < 0 ? b : frame_of_link(W, l0 + k);
    if (k < 0 && kind != 2) { memset(W->pA + 6 * f, 0, 48); memset(W->IAa + 9 * f, 0, 72); memset(W->IAb + 9 * f, 0, 72); memset(W->IAc + 9 * f, 0, 72); continue; }
    double m = mass[f]; const double* I = inertia + 3 * f; const double *w = W->w + 3 * f, *v = W->v + 3 * f;
    double Iw[3] = {I[0] * w[0], I[1] * w[1], I[2] * w[2]}, t[3], fw[3], tw[3];
    double* pA = W->pA + 6 * f;
    v_cross(pA, w, Iw);                                   /* w x I w */
    v_cross(t, w, v); v_scale(pA + 3, t, m);              /* m w x v */
    v_scale(fw, grav, m); v_add(fw, fw, extf + 3 * f);    /* external force (world) incl. gravity */
    mT_vec(t, W->Rw + 9 * f, fw); v_sub(pA + 3, pA + 3, t);
    mT_vec(tw, W->Rw + 9 * f, extt + 3 * f); v_sub(pA, pA, tw);
    double wn = v_len(w), vn = v_len(v);                  /* velocity damping, App. A.2 */
    int dlin = (W->sem & SEM_DAMPING_LINEAR) != 0;   /* damping -m v k instead of -m v (k + k |v|) */
    v_madd(pA, Iw, dlin ? ka : ka + ka * wn);
    double mv[3]; v_scale(mv, v, m); v_madd(pA + 3, mv, dlin ? kl : kl + kl * vn);
    double* A = W->IAa + 9 * f; memset(A, 0, 72); A[0] = I[0]; A[4] = I[1]; A[8] = I[2];
    memset(W->IAb + 9 * f, 0, 72);
    double* C = W->IAc + 9 * f; memset(C, 0, 72); C[0] = C[4] = C[8] = m;
    if (k >= 0) {
      int gl = l0 + k; const int32_t* li = W->link_i + DG_LINK_I_W * gl; double* c = W->c + 6 * gl;
      if (li[3] >= 0) {
        double qdi = qd[li[3]], sa[3], sl[3], t2[3];
        v_scale(sa, W->Sa + 3 * gl, qdi); v_scale(sl, W->Sl + 3 * gl, qdi);
        v_cross(c, w, sa); v_cross(c + 3, w, sl); v_cross(t2, v, sa); v_add(c + 3, c + 3, t2);
      } else memset(c, 0, 48);
    }
  }
  /* pass 2: articulated inertias, leaves to root */
  for (int k = nlb - 1; k >= 0; k--) {
    int gl = l0 + k, f = frame_of_link(W, gl), pf = parent_frame(W, gl);
    const int32_t* li = W->link_i + DG_LINK_I_W * gl;
    double *A = W->IAa + 9 * f, *B = W->IAb + 9 * f, *C = W->IAc + 9 * f, *pA = W->pA + 6 * f, *c = W->c + 6 * gl;
    double Aa[9], Ba[9], Ca[9], pa[6];
    memcpy(Aa, A, 72); memcpy(Ba, B, 72); memcpy(Ca, C, 72);
    double n[3], fo[3];
    ia_mul(A, B, C, c, c + 3, n, fo);
    if (li[3] >= 0) {
      double* U = W->U + 6 * gl; const double *sa = W->Sa + 3 * gl, *sl = W->Sl + 3 * gl;
      ia_mul(A, B, C, sa, sl, U, U + 3);
      double D = v_dot(sa, U) + v_dot(sl, U + 3);
      double tau = jtq[li[3]] + tau_damp[li[3]];
      double uu = tau - (v_dot(sa, pA) + v_dot(sl, pA + 3));
      W->D[gl] = D; W->u[gl] = uu;
      double Dinv = 1.0 / D;
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        Aa[3 * i + j] -= U[i] * U[j] * Dinv; Ba[3 * i + j] -= U[i] * U[3 + j] * Dinv; Ca[3 * i + j] -= U[3 + i] * U[3 + j] * Dinv;
      }
      FL(81);
      ia_mul(Aa, Ba, Ca, c, c + 3, n, fo);
      for (int i = 0; i < 3; i++) { pa[i] = pA[i] + n[i] + U[i] * uu * Dinv; pa[3 + i] = pA[3 + i] + fo[i] + U[3 + i] * uu * Dinv; }
      FL(24);
    } else {
      for (int i = 0; i < 3; i++) { pa[i] = pA[i] + n[i]; pa[3 + i] = pA[3 + i] + fo[i]; }
    }
    /* transform to parent: rotate blocks by E^T (.) E then shift by r */
    const double* E = W->E + 9 * gl; const double* r = W->r + 3 * gl;
    double T1[9], Ar[9], Br[9], Cr[9], K[9];
    mT_mul(T1, E, Aa); m_mul(Ar, T1, E); mT_mul(T1, E, Ba); m_mul(Br, T1, E); mT_mul(T1, E, Ca); m_mul(Cr, T1, E);
    m_skew(K, r);
    double KC[9], BK[9], KBt[9], KCK[9], Brt[9];
    m_mul(KC, K, Cr); m_mul(BK, Br, K); m_T(Brt, Br); m_mul(KBt, K, Brt); m_mul(KCK, KC, K);
    double *Ap = W->IAa + 9 * pf, *Bp = W->IAb + 9 * pf, *Cp = W->IAc + 9 * pf, *pp = W->pA + 6 * pf;
    for (int i = 0; i < 9; i++) { Ap[i] += Ar[i] - BK[i] + KBt[i] - KCK[i]; Bp[i] += Br[i] + KC[i]; Cp[i] += Cr[i]; }
    FL(54);
    double fp[3], np_[3], t[3];
    mT_vec(fp, E, pa + 3); mT_vec(np_, E, pa); v_cross(t, r, fp); v_add(np_, np_, t);
    v_add(pp, pp, np_); v_add(pp + 3, pp + 3, fp);
  }
  /* base acceleration */
  double* a0 = W->acc + 6 * b;
  if (kind == 2) {
    double M[36]; const double *A = W->IAa + 9 * b, *B = W->IAb + 9 * b, *C = W->IAc + 9 * b;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { M[6 * i + j] = A[3 * i + j]; M[6 * i + 3 + j] = B[3 * i + j]; M[6 * (3 + i) + j] = B[3 * j + i]; M[6 * (3 + i) + 3 + j] = C[3 * i + j]; }
    solve_sym(M, 6, W->I0inv + 36 * b);
    const double* Iv = W->I0inv + 36 * b; const double* p0 = W->pA + 6 * b;
    for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s -= Iv[6 * i + j] * p0[j]; a0[i] = s; }
    FL(72);
  } else memset(a0, 0, 48);
  /* pass 3 */
  for (int k = 0; k < nlb; k++) {
    int gl = l0 + k, f = frame_of_link(W, gl), pf = parent_frame(W, gl);
    const int32_t* li = W->link_i + DG_LINK_I_W * gl;
    const double *E = W->E + 9 * gl, *r = W->r + 3 * gl, *ap = W->acc + 6 * pf, *c = W->c + 6 * gl;
    double* a = W->acc + 6 * f; double t[3], t2[3];
    m_vec(a, E, ap); v_cross(t, ap, r); v_add(t2, ap + 3, t); m_vec(a + 3, E, t2);
    v_add(a, a, c); v_add(a + 3, a + 3, c + 3);
    if (li[3] >= 0) {
      const double* U = W->U + 6 * gl;
      double qdd = (W->u[gl] - (v_dot(U, a) + v_dot(U + 3, a + 3))) / W->D[gl];
      W->qdd[li[3]] = qdd;
      v_madd(a, W->Sa + 3 * gl, qdd); v_madd(a + 3, W->Sl + 3 * gl, qdd);
    }
  }
  /* joint reaction wrenches (force_torque_sensor.py:21-23 reads p.getJointState(...)[2]): Bullet's multibody forward
   * dynamics reports  I^A a + p^A  of the child link, in that link's frame, as [Fx Fy Fz Mx My Mz], from the last
   * unconstrained pass - solver impulses are not included  [RECALLED-UNVERIFIED] */
  if (W->need_react) for (int k = 0; k < nlb; k++) {
    int gl = l0 + k, f = frame_of_link(W, gl); double n[3], fo[3]; double* o = ST(W, S_JREACT) + 6 * gl;
    ia_mul(W->IAa + 9 * f, W->IAb + 9 * f, W->IAc + 9 * f, W->acc + 6 * f, W->acc + 6 * f + 3, n, fo);
    for (int i = 0; i < 3; i++) { o[i] = W->pA[6 * f + 3 + i] + fo[i]; o[3 + i] = W->pA[6 * f + i] + n[i]; }
  }
}
/* out = M^-1 gen  for body b using the cached articulated quantities; gen/out are [tau_w(3), f_w(3), joint(nd_b)] */
static void minv_mul(DgoWorld* W, int b, const double* gen, double* out) {
  const int32_t* bi = W->body_i + DG_BODY_I_W * b;
  int kind = bi[0], l0 = bi[1], nlb = bi[2], d0 = bi[3], ndb = bi[4];
  for (int i = 0; i < 6 + ndb; i++) out[i] = 0;
  if (kind == 0) return;
  double p[6 * 65], uu[64], a[6 * 65];  /* index 0 = base, k+1 = link k */
  if (nlb > 64) return;
  memset(p, 0, sizeof(double) * 6 * (size_t)(nlb + 1));
  for (int k = nlb - 1; k >= 0; k--) {
    int gl = l0 + k; const int32_t* li = W->link_i + DG_LINK_I_W * gl;
    int pk = li[1] < 0 ? 0 : li[1] - l0 + 1;
    double pa[6]; memcpy(pa, p + 6 * (k + 1), 48);
    if (li[3] >= 0) {
      const double* U = W->U + 6 * gl;
      double u1 = gen[6 + li[3] - d0] - (v_dot(W->Sa + 3 * gl, pa) + v_dot(W->Sl + 3 * gl, pa + 3));
      uu[k] = u1; double s = u1 / W->D[gl];
      for (int i = 0; i < 6; i++) pa[i] += U[i] * s;
      FL(13);
    }
    const double *E = W->E + 9 * gl, *r = W->r + 3 * gl; double fp[3], np_[3], t[3];
    mT_vec(fp, E, pa + 3); mT_vec(np_, E, pa); v_cross(t, r, fp); v_add(np_, np_, t);
    v_add(p + 6 * pk, p + 6 * pk, np_); v_add(p + 6 * pk + 3, p + 6 * pk + 3, fp);
  }
  if (kind == 2) {
    double rhs[6], t[3];
    mT_vec(t, W->Rw + 9 * b, gen); v_sub(rhs, t, p); mT_vec(t, W->Rw + 9 * b, gen + 3); v_sub(rhs + 3, t, p + 3);
    const double* Iv = W->I0inv + 36 * b;
    for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += Iv[6 * i + j] * rhs[j]; a[i] = s; }
    FL(72);
    m_vec(out, W->Rw + 9 * b, a); m_vec(out + 3, W->Rw + 9 * b, a + 3);
  } else memset(a, 0, 48);
  for (int k = 0; k < nlb; k++) {
    int gl = l0 + k; const int32_t* li = W->link_i + DG_LINK_I_W * gl;
    int pk = li[1] < 0 ? 0 : li[1] - l0 + 1;
    const double *E = W->E + 9 * gl, *r = W->r + 3 * gl, *ap = a + 6 * pk; double* ak = a + 6 * (k + 1); double t[3], t2[3];
    m_vec(ak, E, ap); v_cross(t, ap, r); v_add(t2, ap + 3, t); m_vec(ak + 3, E, t2);
    if (li[3] >= 0) {
      const double* U = W->U + 6 * gl;
      double qdd = (uu[k] - (v_dot(U, ak) + v_dot(U + 3, ak + 3))) / W->D[gl];
      out[6 + li[3] - d0] = qdd;
      v_madd(ak, W->Sa + 3 * gl, qdd); v_madd(ak + 3, W->Sl + 3 * gl, qdd);
    }
  }
}

/* ---------------------------------------------------------------- collision ----------------------------- */
static void shape_world(const DgoWorld* W, int s, double* R, double* p) {
  const int32_t* si = W->shape_i + DG_SHAPE_I_W * s; const double* sf = W->shape_f + DG_SHAPE_F_W * s;
  double Rl[9], t[3]; q_to_mat(Rl, sf + 3); m_mul(R, W->Rw + 9 * si[1], Rl); m_vec(t, W->Rw + 9 * si[1], sf); v_add(p, W->pw + 3 * si[1], t);
}
static void add_contact(DgoWorld* W, int fa, int fb, const double* pa, const double* pb, const double* n, double dist, double mu, double margin, Contact* tmp, int* nt) {
  (void)W;
  if (dist > margin || *nt >= 16) return;
  Contact* c = &tmp[(*nt)++]; c->fa = fa; c->fb = fb; v_cpy(c->pa, pa); v_cpy(c->pb, pb); v_cpy(c->n, n); c->dist = dist; c->mu = mu;
}
/* sphere (centre c, radius r) = A against box B */
static int sphere_box(const double* c, double r, const double* Rb, const double* pb, const double* h, double* pa_out, double* pb_out, double* n, double* dist) {
  double d[3], cl[3], q[3]; v_sub(d, c, pb); mT_vec(cl, Rb, d);
  int inside = 1;
  for (int i = 0; i < 3; i++) { q[i] = cl[i]; if (q[i] > h[i]) { q[i] = h[i]; inside = 0; } else if (q[i] < -h[i]) { q[i] = -h[i]; inside = 0; } }
  double nl[3] = {0, 0, 0};
  if (!inside) {
    double dv[3]; v_sub(dv, cl, q); double len = v_len(dv);
    if (len < 1e-12) return 0;
    v_scale(nl, dv, 1.0 / len); *dist = len - r;
  } else {
    int ax = 0; double best = 1e300;
    for (int i = 0; i < 3; i++) { double pen = h[i] - fabs(cl[i]); if (pen < best) { best = pen; ax = i; } }
    nl[ax] = cl[ax] >= 0 ? 1.0 : -1.0; q[ax] = nl[ax] * h[ax]; *dist = -best - r;
  }
  m_vec(n, Rb, nl); double t[3]; m_vec(t, Rb, q); v_add(pb_out, pb, t);
  v_scale(t, n, -r); v_add(pa_out, c, t);
  return 1;
}
/* point p = A against box B with margin: contact if inside the margin-inflated box */
static int point_box(const double* p, const double* Rb, const double* pb, const double* h, double margin, double* pb_out, double* n, double* dist) {
  double d[3], pl[3]; v_sub(d, p, pb); mT_vec(pl, Rb, d);
  int ax = 0; double best = 1e300;
  for (int i = 0; i < 3; i++) { double pen = h[i] - fabs(pl[i]); if (pen < -margin) return 0; if (pen < best) { best = pen; ax = i; } }
  double nl[3] = {0, 0, 0}, q[3] = {pl[0], pl[1], pl[2]};
  nl[ax] = pl[ax] >= 0 ? 1.0 : -1.0; q[ax] = nl[ax] * h[ax];
  m_vec(n, Rb, nl); double t[3]; m_vec(t, Rb, q); v_add(pb_out, pb, t); *dist = -best;
  return 1;
}
static void seg_closest(const double* p1, const double* d1, const double* p2, const double* d2, double* s, double* t) {
  /* closest parameters on segments p1 + s d1, p2 + t d2, s,t in [0,1] */
  double r[3]; v_sub(r, p1, p2);
  double a = v_dot(d1, d1), e = v_dot(d2, d2), f = v_dot(d2, r);
  if (a <= 1e-18 && e <= 1e-18) { *s = *t = 0; return; }
  if (a <= 1e-18) { *s = 0; *t = fmin(fmax(f / e, 0), 1); return; }
  double c = v_dot(d1, r);
  if (e <= 1e-18) { *t = 0; *s = fmin(fmax(-c / a, 0), 1); return; }
  double b = v_dot(d1, d2), den = a * e - b * b;
  *s = den > 1e-18 ? fmin(fmax((b * f - c * e) / den, 0), 1) : 0.0;
  *t = (b * (*s) + f) / e;
  if (*t < 0) { *t = 0; *s = fmin(fmax(-c / a, 0), 1); } else if (*t > 1) { *t = 1; *s = fmin(fmax((b - c) / a, 0), 1); }
}
static int sphere_sphere(const double* c1, double r1, const double* c2, double r2, double* pa, double* pb, double* n, double* dist) {
  double d[3]; v_sub(d, c1, c2); double len = v_len(d);
  if (len < 1e-12) { v_set(n, 0, 0, 1); } else v_scale(n, d, 1.0 / len);
  *dist = len - r1 - r2; double t[3]; v_scale(t, n, -r1); v_add(pa, c1, t); v_scale(t, n, r2); v_add(pb, c2, t);
  return 1;
}
/* rounded view of a shape for non-box pairs: segment endpoints + radius (sphere: degenerate segment) */
static void as_capsule(int type, const double* dims, const double* R, const double* p, double* e0, double* e1, double* rad) {
  double half = 0; *rad = dims[0];
  if (type == SHAPE_CAPSULE) half = dims[1];
  else if (type == SHAPE_CYLINDER) half = dims[1] - dims[0] > 0 ? dims[1] - dims[0] : 0;
  double ax[3] = {R[2] * half, R[5] * half, R[8] * half};
  v_sub(e0, p, ax); v_add(e1, p, ax);
}
/* ---- reduced convex hulls: mesh links (diy_gym/model.py:65 loads them as convex hulls; here at most 32 vertices, compiler/mesh.py).
 * Signed distance of world point p to the hull = the largest plane distance; inside or within `margin`: the outward normal of that
 * plane, the point moved onto it, and the (negative) distance. */
static int point_hull(const double* p, const double* R, const double* pos, const double* planes, int np_, double margin, double* surf, double* n, double* dist) {
  double d[3], pl[3]; v_sub(d, p, pos); mT_vec(pl, R, d);
  double best = -1e300; int bi = 0;
  for (int k = 0; k < np_; k++) { const double* q = planes + 4 * k; double s = q[0] * pl[0] + q[1] * pl[1] + q[2] * pl[2] - q[3]; if (s > best) { best = s; bi = k; } }
  if (best > margin) return 0;
  m_vec(n, R, planes + 4 * bi); *dist = best;
  for (int i = 0; i < 3; i++) surf[i] = p[i] - n[i] * best;
  return 1;
}
/* convex pair with at least one hull (the other a hull or a box): vertices of each inside the other (no edge-edge contacts, as in the
 * box-box routine); contact convention (pa on A, pb on B, normal from B towards A) */
static void collide_convex(DgoWorld* W, const int32_t* ia, const double* fa, const double* Ra, const double* pa, const int32_t* ib, const double* fb,
                           const double* Rb, const double* pb, double mu, double margin, Contact* loc, int* nloc) {
  const double* H = W->hull_f;
  for (int side = 0; side < 2; side++) {
    const int32_t *ix = side ? ib : ia, *iy = side ? ia : ib; const double *fx = side ? fb : fa, *fy = side ? fa : fb;
    const double *Rx = side ? Rb : Ra, *px = side ? pb : pa, *Ry = side ? Ra : Rb, *py = side ? pa : pb;
    int nvx = ix[5] > 0 ? ix[5] : 8;
    for (int k = 0; k < nvx; k++) {
      double vl[3], t[3], pt[3], surf[3], n[3], dist;
      if (ix[5] > 0) { const double* v = H + ix[4] + 3 * k; v_cpy(vl, v); }
      else { vl[0] = (k & 1 ? 1 : -1) * fx[7]; vl[1] = (k & 2 ? 1 : -1) * fx[8]; vl[2] = (k & 4 ? 1 : -1) * fx[9]; }
      m_vec(t, Rx, vl); v_add(pt, px, t);
      int hit = iy[5] > 0 ? point_hull(pt, Ry, py, H + iy[6], iy[7], margin, surf, n, &dist) : point_box(pt, Ry, py, fy + 7, margin, surf, n, &dist);
      if (!hit) continue;
      if (side == 0) add_contact(W, ia[1], ib[1], pt, surf, n, dist, mu, margin, loc, nloc);
      else { double nn[3]; v_scale(nn, n, -1.0); add_contact(W, ia[1], ib[1], surf, pt, nn, dist, mu, margin, loc, nloc); }
    }
  }
}
static void collide_pair(DgoWorld* W, int sa, int sb, double margin) {
  const int32_t *ia = W->shape_i + DG_SHAPE_I_W * sa, *ib = W->shape_i + DG_SHAPE_I_W * sb;
  const double *fa = W->shape_f + DG_SHAPE_F_W * sa, *fb = W->shape_f + DG_SHAPE_F_W * sb;
  double Ra[9], pa[3], Rb[9], pb[3];
  shape_world(W, sa, Ra, pa); shape_world(W, sb, Rb, pb);
  double dc[3]; v_sub(dc, pa, pb);
  double reach = fa[11] + fb[11] + margin;
  if (v_dot(dc, dc) > reach * reach) return;
  int ta = ia[2], tb = ib[2];
  double mu = PR(W, P_FRICTION)[sa] * PR(W, P_FRICTION)[sb];
  Contact tmp[16]; int nt = 0;
  double ca[3], cb[3], n[3], dist;
  int ha = ia[5] > 0, hb_ = ib[5] > 0;
  if ((ha && (hb_ || tb == SHAPE_BOX)) || (hb_ && (ha || ta == SHAPE_BOX))) {
    /* mesh links: reduced hull against a box or another hull (spheres / capsules / cylinders meet the fitted proxy below) */
    Contact loc[16]; int nloc = 0;
    collide_convex(W, ia, fa, Ra, pa, ib, fb, Rb, pb, mu, margin, loc, &nloc);
    for (int i = 0; i < nloc; i++) for (int j = i + 1; j < nloc; j++) if (loc[j].dist < loc[i].dist) { Contact t = loc[i]; loc[i] = loc[j]; loc[j] = t; }
    if (nloc > 4) nloc = 4;
    for (int i = 0; i < nloc; i++) tmp[nt++] = loc[i];
  } else if (ta != SHAPE_BOX && tb != SHAPE_BOX) {
    double a0[3], a1[3], b0[3], b1[3], ra, rb, d1[3], d2[3], s, t, c1[3], c2[3];
    as_capsule(ta, fa + 7, Ra, pa, a0, a1, &ra); as_capsule(tb, fb + 7, Rb, pb, b0, b1, &rb);
    v_sub(d1, a1, a0); v_sub(d2, b1, b0); seg_closest(a0, d1, b0, d2, &s, &t);
    v_cpy(c1, a0); v_madd(c1, d1, s); v_cpy(c2, b0); v_madd(c2, d2, t);
    sphere_sphere(c1, ra, c2, rb, ca, cb, n, &dist);
    add_contact(W, ia[1], ib[1], ca, cb, n, dist, mu, margin, tmp, &nt);
  } else {
    /* at least one box: make B the box (flip at the end if we swapped) */
    int swap = (tb != SHAPE_BOX);
    const double *Rx = swap ? Rb : Ra, *px = swap ? pb : pa, *dx = swap ? fb + 7 : fa + 7;   /* the non-reference shape X */
    const double *Rbx = swap ? Ra : Rb, *pbx = swap ? pa : pb, *hb = swap ? fa + 7 : fb + 7; /* the reference box */
    int tx = swap ? tb : ta; int fx = swap ? ib[1] : ia[1], fbx = swap ? ia[1] : ib[1];
    Contact loc[16]; int nloc = 0;
    if (tx == SHAPE_SPHERE) {
      if (sphere_box(px, dx[0], Rbx, pbx, hb, ca, cb, n, &dist)) add_contact(W, fx, fbx, ca, cb, n, dist, mu, margin, loc, &nloc);
    } else if (tx == SHAPE_CAPSULE) {
      double e0[3], e1[3], rad, mid[3], dseg[3], rel[3];
      as_capsule(tx, dx, Rx, px, e0, e1, &rad);
      v_sub(dseg, e1, e0); v_sub(rel, pbx, e0);
      double dd = v_dot(dseg, dseg), tt = dd > 1e-18 ? fmin(fmax(v_dot(rel, dseg) / dd, 0), 1) : 0.0;
      v_cpy(mid, e0); v_madd(mid, dseg, tt);
      if (sphere_box(e0, rad, Rbx, pbx, hb, ca, cb, n, &dist)) add_contact(W, fx, fbx, ca, cb, n, dist, mu, margin, loc, &nloc);
      if (dd > 1e-18 && sphere_box(e1, rad, Rbx, pbx, hb, ca, cb, n, &dist)) add_contact(W, fx, fbx, ca, cb, n, dist, mu, margin, loc, &nloc);
      if (tt > 1e-6 && tt < 1 - 1e-6 && sphere_box(mid, rad, Rbx, pbx, hb, ca, cb, n, &dist)) add_contact(W, fx, fbx, ca, cb, n, dist, mu, margin, loc, &nloc);
    } else if (tx == SHAPE_CYLINDER) {
      double zc[3] = {Rx[2], Rx[5], Rx[8]}, xc[3] = {Rx[0], Rx[3], Rx[6]}, yc[3] = {Rx[1], Rx[4], Rx[7]};
      for (int cap = -1; cap <= 1; cap += 2) {
        double cc[3]; v_cpy(cc, px); v_madd(cc, zc, cap * dx[1]);
        for (int k = 0; k < 6; k++) {
          double nk[3] = {-Rbx[k % 3] * (k < 3 ? 1 : -1), -Rbx[3 + k % 3] * (k < 3 ? 1 : -1), -Rbx[6 + k % 3] * (k < 3 ? 1 : -1)}; /* -face normal */
          double t[3]; v_cpy(t, nk); v_madd(t, zc, -v_dot(nk, zc));
          double len = v_len(t), pt[3];
          int npt = 1;
          if (len < 1e-6) npt = 4;  /* cap parallel to this face: take four rim points instead */
          for (int j4 = 0; j4 < npt; j4++) {
            if (npt == 4) { const double* bx = (j4 & 1) ? yc : xc; double sg = (j4 & 2) ? -1.0 : 1.0; v_cpy(pt, cc); v_madd(pt, bx, sg * dx[0]); }
            else { v_cpy(pt, cc); v_madd(pt, t, dx[0] / len); }
            if (point_box(pt, Rbx, pbx, hb, margin, cb, n, &dist)) {
              int dup = 0;
              for (int j = 0; j < nloc; j++) { double dd[3]; v_sub(dd, loc[j].pa, pt); if (v_dot(dd, dd) < 1e-12) dup = 1; }
              if (!dup) add_contact(W, fx, fbx, pt, cb, n, dist, mu, margin, loc, &nloc);
            }
          }
        }
      }
    } else { /* box X against box B: corners of X in B, then corners of B in X */
      for (int k = 0; k < 8; k++) {
        double cl[3] = {(k & 1 ? 1 : -1) * dx[0], (k & 2 ? 1 : -1) * dx[1], (k & 4 ? 1 : -1) * dx[2]}, pt[3], t[3];
        m_vec(t, Rx, cl); v_add(pt, px, t);
        if (point_box(pt, Rbx, pbx, hb, margin, cb, n, &dist)) add_contact(W, fx, fbx, pt, cb, n, dist, mu, margin, loc, &nloc);
      }
      for (int k = 0; k < 8; k++) {
        double cl[3] = {(k & 1 ? 1 : -1) * hb[0], (k & 2 ? 1 : -1) * hb[1], (k & 4 ? 1 : -1) * hb[2]}, pt[3], t[3], nn[3], cx[3];
        m_vec(t, Rbx, cl); v_add(pt, pbx, t);
        if (point_box(pt, Rx, px, dx, margin, cx, nn, &dist)) { v_scale(nn, nn, -1.0); add_contact(W, fx, fbx, cx, pt, nn, dist, mu, margin, loc, &nloc); }
      }
    }
    /* keep the 4 deepest */
    for (int i = 0; i < nloc; i++) for (int j = i + 1; j < nloc; j++) if (loc[j].dist < loc[i].dist) { Contact t = loc[i]; loc[i] = loc[j]; loc[j] = t; }
    if (nloc > 4) nloc = 4;
    for (int i = 0; i < nloc; i++) {
      Contact c = loc[i];
      if (swap) { Contact s2 = c; s2.fa = c.fb; s2.fb = c.fa; v_cpy(s2.pa, c.pb); v_cpy(s2.pb, c.pa); v_scale(s2.n, c.n, -1.0); c = s2; }
      tmp[nt++] = c;
    }
  }
  /* capacity max_contacts (an extension key of the YAML, default 16 with a floating body, else 8): when the list is full a new
     contact replaces the SHALLOWEST stored one if it is deeper, so that what gets lost is a grazing contact, never the wall the
     robot is pushing into; every loss is counted (dgo_contacts_dropped).  pybullet has no such cap. */
  for (int i = 0; i < nt; i++) {
    if (W->ncontacts < W->maxc) { W->contacts[W->ncontacts++] = tmp[i]; continue; }
    W->dropped++;
    int worst = 0;
    for (int j = 1; j < W->ncontacts; j++) if (W->contacts[j].dist > W->contacts[worst].dist) worst = j;
    if (W->ncontacts > 0 && tmp[i].dist < W->contacts[worst].dist) W->contacts[worst] = tmp[i];
  }
}
static void collide(DgoWorld* W) {
  W->ncontacts = 0;
  double margin = HF(W, contact_margin);
  for (int k = 0; k < W->npair; k++) collide_pair(W, W->pair_i[2 * k], W->pair_i[2 * k + 1], margin);
}

/* ---------------------------------------------------------------- constraint rows + PGS ----------------- */
static int body_of_frame(const DgoWorld* W, int f) { return f < W->nb ? f : W->link_i[DG_LINK_I_W * (f - W->nb)]; }
/* Jacobian (generalized force per unit force along `dir` at world point `p` on frame f) */
static void point_jacobian(DgoWorld* W, int f, const double* p, const double* dir, double* J) {
  int b = body_of_frame(W, f); const int32_t* bi = W->body_i + DG_BODY_I_W * b;
  int ndb = bi[4], d0 = bi[3];
  for (int i = 0; i < 6 + ndb; i++) J[i] = 0;
  if (bi[0] == 0) return;
  if (bi[0] == 2) { double rel[3]; v_sub(rel, p, W->pw + 3 * b); v_cross(J, rel, dir); v_cpy(J + 3, dir); }
  int gl = f < W->nb ? -1 : f - W->nb;
  while (gl >= 0) {
    const int32_t* li = W->link_i + DG_LINK_I_W * gl; const double* lf = W->link_f + DG_LINK_F_W * gl; int fr = frame_of_link(W, gl);
    if (li[2] == 1) {
      double aw[3], dw[3], o[3], rel[3], t[3];
      m_vec(aw, W->Rw + 9 * fr, lf + 10); m_vec(dw, W->Rw + 9 * fr, lf + 7); v_sub(o, W->pw + 3 * fr, dw);
      v_sub(rel, p, o); v_cross(t, aw, rel); J[6 + li[3] - d0] = v_dot(dir, t);
    } else if (li[2] == 2) { double aw[3]; m_vec(aw, W->Rw + 9 * fr, lf + 10); J[6 + li[3] - d0] = v_dot(dir, aw); }
    gl = li[1];
  }
}
/* generalized force per unit TORQUE along `dir` on frame f (same coordinates as point_jacobian) */
static void angular_jacobian(DgoWorld* W, int f, const double* dir, double* J) {
  int b = body_of_frame(W, f); const int32_t* bi = W->body_i + DG_BODY_I_W * b;
  int ndb = bi[4], d0 = bi[3];
  for (int i = 0; i < 6 + ndb; i++) J[i] = 0;
  if (bi[0] == 0) return;
  if (bi[0] == 2) v_cpy(J, dir);
  for (int gl = f < W->nb ? -1 : f - W->nb; gl >= 0; gl = W->link_i[DG_LINK_I_W * gl + 1]) {
    const int32_t* li = W->link_i + DG_LINK_I_W * gl; const double* lf = W->link_f + DG_LINK_F_W * gl;
    if (li[2] == 1) { double aw[3]; m_vec(aw, W->Rw + 9 * frame_of_link(W, gl), lf + 10); J[6 + li[3] - d0] = v_dot(dir, aw); }
  }
}
static void body_genvel(const DgoWorld* W, int b, double* gv) {
  const int32_t* bi = W->body_i + DG_BODY_I_W * b;
  v_cpy(gv, ST(W, S_BOMEGA) + 3 * b); v_cpy(gv + 3, ST(W, S_BVEL) + 3 * b);
  for (int i = 0; i < bi[4]; i++) gv[6 + i] = ST(W, S_QD)[bi[3] + i];
}
static double dotn(const double* a, const double* b, int n) { double s = 0; for (int i = 0; i < n; i++) s += a[i] * b[i]; FL(2 * n); return s; }
static int body_ndof(const DgoWorld* W, int b) { return 6 + W->body_i[DG_BODY_I_W * b + 4]; }

static Row* new_row(DgoWorld* W) { Row* r = &W->rows[W->nrows++]; memset(r, 0, sizeof(Row)); r->bodyA = r->bodyB = -1; r->parent = -1; r->motor_dof = -1; r->unit_dof = -1; return r; }
/* finish a row: M = Minv J, diag, relative velocity */
static double finish_row(DgoWorld* W, Row* r) {
  double den = 0, rel = 0, gv[MAXROWDOF];
  if (r->bodyA >= 0) { int n = body_ndof(W, r->bodyA); minv_mul(W, r->bodyA, r->JA, r->MA); den += dotn(r->JA, r->MA, n); body_genvel(W, r->bodyA, gv); rel += dotn(r->JA, gv, n); }
  if (r->bodyB >= 0) { int n = body_ndof(W, r->bodyB); minv_mul(W, r->bodyB, r->JB, r->MB); den += dotn(r->JB, r->MB, n); body_genvel(W, r->bodyB, gv); rel += dotn(r->JB, gv, n); }
  r->diag_inv = den > 1e-30 ? 1.0 / den : 0.0;
  return rel;
}
static void plane_space(const double* n, double* p, double* q) {
  if (fabs(n[2]) > 0.7071067811865475244) { double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a); p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k; q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1]; }
  else { double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a); p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0; q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k; }
}
static void build_rows(DgoWorld* W, double h) {
  W->nrows = 0;
  const double *q = ST(W, S_Q), *qd = ST(W, S_QD);
  double dt = HF(W, dt), erp = HF(W, erp);
  for (int b = 0; b < W->nb; b++) {
    const int32_t* bi = W->body_i + DG_BODY_I_W * b;
    if (bi[0] == 0) continue;
    int l0 = bi[1], nlb = bi[2], d0 = bi[3];
    /* joint limit rows: only when violated (App. A.1 / A.2) */
    for (int k = 0; k < nlb; k++) {
      const int32_t* li = W->link_i + DG_LINK_I_W * (l0 + k); const double* lf = W->link_f + DG_LINK_F_W * (l0 + k);
      if (li[3] < 0 || !li[4]) continue;
      for (int side = 0; side < 2; side++) {
        double pen = side == 0 ? q[li[3]] - lf[20] : lf[21] - q[li[3]];
        if (pen > 0) continue;
        Row* r = new_row(W); r->bodyA = b; r->unit_dof = li[3] - d0;
        r->JA[6 + li[3] - d0] = side == 0 ? 1.0 : -1.0;
        double rel = finish_row(W, r);
        r->rhs = (-rel + (-pen) * erp / h) * r->diag_inv; r->lo = 0; r->hi = HF(W, limit_max_impulse);
      }
    }
    /* motor rows (App. A.3) */
    for (int k = 0; k < nlb; k++) {
      const int32_t* li = W->link_i + DG_LINK_I_W * (l0 + k);
      int d = li[3]; if (d < 0) continue;
      double maxf = ST(W, S_MMAXF)[d];
      ST(W, S_MAPPLIED)[d] = 0;
      if (maxf <= 0) continue;
      Row* r = new_row(W); r->bodyA = b; r->unit_dof = d - d0; r->motor_dof = d; r->JA[6 + d - d0] = 1.0;
      double rel = finish_row(W, r);
      double kp = ST(W, S_MKP)[d], kd = ST(W, S_MKD)[d], tp = ST(W, S_MTPOS)[d], tv = ST(W, S_MTVEL)[d];
      double desired = kp * (tp - q[d]) / h + qd[d] + kd * (tv - qd[d]);
      r->rhs = (desired - rel) * r->diag_inv; { double cdt = (W->sem & SEM_MOTOR_CLAMP_SUBSTEP) ? h : dt; r->lo = -maxf * cdt; r->hi = maxf * cdt; }
    }
  }
  /* fixed constraints between models (diy_gym/model.py:69-77, p.createConstraint(..., JOINT_FIXED, ...)): the joint frame on
   * the parent link (pos_a, quat_a in its COM frame) and the one on the child link are held together by three point rows
   * along the world axes and three angular rows, error reduction erp per sub-step, impulse bound max_force x dt.
   * [RECALLED-UNVERIFIED: Bullet's btMultiBodyFixedConstraint builds its rows in the parent frame; same solution set] */
  for (int k = 0; k < W->ncons; k++) {
    const int32_t* ci = W->cons_i + DG_CONS_I_W * k; const double* cf = W->cons_f + DG_CONS_F_W * k;
    int fa = ci[0], fb = ci[1], ba = body_of_frame(W, fa), bb = body_of_frame(W, fb);
    double Pa[3], Pb[3], t[3], qa[4], qb[4], qaw[4], qbw[4], qbi[4], dq[4], th[3];
    m_vec(t, W->Rw + 9 * fa, cf); v_add(Pa, W->pw + 3 * fa, t);
    m_vec(t, W->Rw + 9 * fb, cf + 7); v_add(Pb, W->pw + 3 * fb, t);
    mat_to_q(qa, W->Rw + 9 * fa); mat_to_q(qb, W->Rw + 9 * fb);
    q_mul(qaw, qa, cf + 3); q_mul(qbw, qb, cf + 10);
    qbi[0] = -qbw[0]; qbi[1] = -qbw[1]; qbi[2] = -qbw[2]; qbi[3] = qbw[3];
    q_mul(dq, qaw, qbi);                                   /* rotation that takes the child joint frame to the parent one */
    { double vn = v_len(dq), ang = 2 * atan2(vn, dq[3]); if (ang > M_PI) ang -= 2 * M_PI;
      if (vn < 1e-12) v_set(th, 0, 0, 0); else v_scale(th, dq, ang / vn); }
    for (int i = 0; i < 6; i++) {
      double e[3] = {0, 0, 0}, ne[3] = {0, 0, 0}; e[i % 3] = 1; ne[i % 3] = -1;
      Row* r = new_row(W);
      if (W->body_i[DG_BODY_I_W * ba] != 0) { r->bodyA = ba; if (i < 3) point_jacobian(W, fa, Pa, e, r->JA); else angular_jacobian(W, fa, e, r->JA); }
      if (W->body_i[DG_BODY_I_W * bb] != 0) { r->bodyB = bb; if (i < 3) point_jacobian(W, fb, Pb, ne, r->JB); else angular_jacobian(W, fb, ne, r->JB); }
      double rel = finish_row(W, r);
      double err = i < 3 ? Pa[i] - Pb[i] : th[i - 3];
      r->rhs = (-rel - err * erp / h) * r->diag_inv; r->lo = -cf[14] * dt; r->hi = cf[14] * dt;
    }
  }
  /* contacts: all normals first, then friction rows */
  int first_contact = W->nrows;
  double cerp = HF(W, contact_erp), slop = HF(W, linear_slop);
  for (int k = 0; k < W->ncontacts; k++) {
    Contact* c = &W->contacts[k]; Row* r = new_row(W);
    int ba = body_of_frame(W, c->fa), bb = body_of_frame(W, c->fb);
    if (W->body_i[DG_BODY_I_W * ba] != 0) { r->bodyA = ba; point_jacobian(W, c->fa, c->pa, c->n, r->JA); }
    if (W->body_i[DG_BODY_I_W * bb] != 0) { double nn[3]; v_scale(nn, c->n, -1.0); r->bodyB = bb; point_jacobian(W, c->fb, c->pb, nn, r->JB); }
    double rel = finish_row(W, r);
    double pen = c->dist + slop, pos_err = 0, vel_err = -rel;
    if (pen > 0) vel_err -= pen / h; else pos_err = -pen * cerp / h;
    r->rhs = (pos_err + vel_err) * r->diag_inv; r->lo = 0; r->hi = 1e10; r->mu = c->mu;
  }
  for (int k = 0; k < W->ncontacts; k++) {
    Contact* c = &W->contacts[k];
    double t1[3], t2[3]; plane_space(c->n, t1, t2);
    for (int dir = 0; dir < 2; dir++) {
      const double* t = dir == 0 ? t1 : t2; Row* r = new_row(W); r->parent = first_contact + k; r->mu = c->mu;
      int ba = body_of_frame(W, c->fa), bb = body_of_frame(W, c->fb);
      if (W->body_i[DG_BODY_I_W * ba] != 0) { r->bodyA = ba; point_jacobian(W, c->fa, c->pa, t, r->JA); }
      if (W->body_i[DG_BODY_I_W * bb] != 0) { double nn[3]; v_scale(nn, t, -1.0); r->bodyB = bb; point_jacobian(W, c->fb, c->pb, nn, r->JB); }
      double rel = finish_row(W, r);
      r->rhs = -rel * r->diag_inv; r->lo = 0; r->hi = 0;
    }
  }
}
static void solve_row(DgoWorld* W, Row* r) {
  double d = r->rhs;
  if (r->bodyA >= 0) d -= dotn(r->JA, W->dv + MAXROWDOF * r->bodyA, body_ndof(W, r->bodyA)) * r->diag_inv;
  if (r->bodyB >= 0) d -= dotn(r->JB, W->dv + MAXROWDOF * r->bodyB, body_ndof(W, r->bodyB)) * r->diag_inv;
  double lo = r->lo, hi = r->hi;
  if (r->parent >= 0) { double nimp = W->rows[r->parent].applied; hi = r->mu * nimp; lo = -hi; }
  double sum = r->applied + d;
  if (sum < lo) { d = lo - r->applied; sum = lo; } else if (sum > hi) { d = hi - r->applied; sum = hi; }
  r->applied = sum; FL(6);
  if (r->bodyA >= 0) { double* dv = W->dv + MAXROWDOF * r->bodyA; int n = body_ndof(W, r->bodyA); for (int i = 0; i < n; i++) dv[i] += r->MA[i] * d; FL(2 * n); }
  if (r->bodyB >= 0) { double* dv = W->dv + MAXROWDOF * r->bodyB; int n = body_ndof(W, r->bodyB); for (int i = 0; i < n; i++) dv[i] += r->MB[i] * d; FL(2 * n); }
}
static void pgs(DgoWorld* W) {
  memset(W->dv, 0, sizeof(double) * MAXROWDOF * (size_t)W->nb);
  int nnc = 0; while (nnc < W->nrows && W->rows[nnc].parent < 0 && (W->rows[nnc].unit_dof >= 0)) nnc++;
  for (int it = 0; it < W->iters; it++) {
    for (int j = 0; j < nnc; j++) solve_row(W, &W->rows[(it & 1) ? j : nnc - 1 - j]);
    for (int j = nnc; j < W->nrows; j++) solve_row(W, &W->rows[j]);
  }
  for (int i = 0; i < W->nrows; i++) if (W->rows[i].motor_dof >= 0) ST(W, S_MAPPLIED)[W->rows[i].motor_dof] = W->rows[i].applied;
}

/* ---------------------------------------------------------------- one stepSimulation -------------------- */
static void integrate_base_quat(double* q, const double* om, double h) {
  double ang = v_len(om), ax[3];
  if (ang * h > 0.25 * M_PI) ang = 0.25 * M_PI / h;   /* angular motion threshold */
  if (ang < 0.001) v_scale(ax, om, 0.5 * h - h * h * h * 0.020833333333 * ang * ang); else v_scale(ax, om, sin(0.5 * ang * h) / ang);
  double dq[4] = {ax[0], ax[1], ax[2], cos(0.5 * ang * h)}, out[4];
  q_mul(out, dq, q); q_norm(out); memcpy(q, out, 32);
}
/* restates p.stepSimulation() as configured at diy_gym/diy_gym.py:76-82 */
void dgo_step_physics(DgoWorld* W) {
  double dt = HF(W, dt), h = dt / W->substeps;
  double *qd = ST(W, S_QD), *q = ST(W, S_Q);
  double tau_damp[256];
  for (int d = 0; d < W->nd && d < 256; d++) tau_damp[d] = -PR(W, P_JDAMP)[d] * qd[d];  /* once per outer step */
  for (int sub = 0; sub < W->substeps; sub++) {
    dgo_forward_kinematics(W); velocities(W);
    for (int b = 0; b < W->nb; b++) {
      const int32_t* bi = W->body_i + DG_BODY_I_W * b;
      if (bi[0] == 0) continue;
      aba_body(W, b, tau_damp);
      if (bi[0] == 2) {  /* spatial -> classical base acceleration, world frame */
        double *a0 = W->acc + 6 * b, t[3], lin[3], aw[3];
        v_cross(t, W->w + 3 * b, W->v + 3 * b); v_add(lin, a0 + 3, t);
        m_vec(aw, W->Rw + 9 * b, a0); v_madd(ST(W, S_BOMEGA) + 3 * b, aw, h);
        m_vec(aw, W->Rw + 9 * b, lin); v_madd(ST(W, S_BVEL) + 3 * b, aw, h);
      }
      for (int i = 0; i < bi[4]; i++) qd[bi[3] + i] += h * W->qdd[bi[3] + i];
    }
    collide(W);
    build_rows(W, h);
    pgs(W);
    double maxv = HF(W, max_joint_vel);
    for (int b = 0; b < W->nb; b++) {
      const int32_t* bi = W->body_i + DG_BODY_I_W * b; const double* dv = W->dv + MAXROWDOF * b;
      if (bi[0] == 0) continue;
      if (bi[0] == 2) {
        for (int i = 0; i < 3; i++) { ST(W, S_BOMEGA)[3 * b + i] += dv[i]; ST(W, S_BVEL)[3 * b + i] += dv[3 + i]; }
        for (int i = 0; i < 3; i++) ST(W, S_BPOS)[3 * b + i] += h * ST(W, S_BVEL)[3 * b + i];
        integrate_base_quat(ST(W, S_BQUAT) + 4 * b, ST(W, S_BOMEGA) + 3 * b, h);
      }
      for (int i = 0; i < bi[4]; i++) {
        int d = bi[3] + i; qd[d] += dv[6 + i];
        if (qd[d] > maxv) qd[d] = maxv; else if (qd[d] < -maxv) qd[d] = -maxv;
        q[d] += h * qd[d];
      }
    }
    if ((W->sem & SEM_WRENCH_FIRST_SUBSTEP) && sub == 0) {   /* applied wrenches / joint torques act during the first internal sub-step only */
      memset(ST(W, S_EXTF), 0, sizeof(double) * 3 * (size_t)W->nframes);
      memset(ST(W, S_EXTT), 0, sizeof(double) * 3 * (size_t)W->nframes);
      memset(ST(W, S_JTORQUE), 0, sizeof(double) * (size_t)W->nd);
    }
  }
  write_link_cache(W);
  memset(ST(W, S_EXTF), 0, sizeof(double) * 3 * (size_t)W->nframes);
  memset(ST(W, S_EXTT), 0, sizeof(double) * 3 * (size_t)W->nframes);
  memset(ST(W, S_JTORQUE), 0, sizeof(double) * (size_t)W->nd);
}

/* ---------------------------------------------------------------- frame queries ------------------------- */
/* COM-frame pose and velocity of a frame as cached after the last step (getBasePositionAndOrientation / getLinkState[0,1,6,7]) */
static void frame_com_state(const DgoWorld* W, int f, double* pos, double* quat, double* vel, double* om) {
  if (f < W->nb) { v_cpy(pos, ST(W, S_BPOS) + 3 * f); memcpy(quat, ST(W, S_BQUAT) + 4 * f, 32); v_cpy(vel, ST(W, S_BVEL) + 3 * f); v_cpy(om, ST(W, S_BOMEGA) + 3 * f); }
  else { int gl = f - W->nb; v_cpy(pos, ST(W, S_LPOS) + 3 * gl); memcpy(quat, ST(W, S_LQUAT) + 4 * gl, 32); v_cpy(vel, ST(W, S_LVEL) + 3 * gl); v_cpy(om, ST(W, S_LOMEGA) + 3 * gl); }
}
/* URDF link-frame pose (getLinkState[4,5]); for a base frame this is the COM pose, as reach_target.py:21-30 uses it */
static void frame_link_pose(const DgoWorld* W, int f, double* pos, double* quat) {
  double v[3], o[3];
  frame_com_state(W, f, pos, quat, v, o);
  if (f >= W->nb) {
    const double* lf = W->link_f + DG_LINK_F_W * (f - W->nb);
    double R[9], t[3], qi[4] = {-lf[16], -lf[17], -lf[18], lf[19]}, qo[4];
    q_to_mat(R, quat); m_vec(t, R, lf + 7); v_sub(pos, pos, t);
    q_mul(qo, quat, qi); memcpy(quat, qo, 32);
  }
}
void dgo_frame_state(DgoWorld* W, int f, double* out) {  /* com pos3 quat4 vel3 om3, link pos3 quat4 */
  frame_com_state(W, f, out, out + 3, out + 7, out + 10); frame_link_pose(W, f, out + 13, out + 16);
}

/* ---------------------------------------------------------------- inverse kinematics -------------------- */
/* restates p.calculateInverseKinematics as called at diy_gym/addons/controllers/ik_controller.py:61-69:
 * damped least squares on the end-effector link frame in base coordinates over all DoF of the body, <= ik_iters
 * iterations while the position residual exceeds ik_threshold, optional null-space term.  out has nd_b entries. */
static void ik_fk(const DgoWorld* W, int b, int ee_gl, const double* qb, double* pos, double* R, double* Jl, double* Ja) {
  const int32_t* bi = W->body_i + DG_BODY_I_W * b; int l0 = bi[1], d0 = bi[3], ndb = bi[4];
  int chain[64], nc = 0;
  for (int gl = ee_gl; gl >= 0; gl = W->link_i[DG_LINK_I_W * gl + 1]) chain[nc++] = gl;
  double Rc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pc[3] = {0, 0, 0};
  double ax_w[64 * 3], org_w[64 * 3]; int jt[64], jd[64];
  (void)l0;
  for (int i = nc - 1; i >= 0; i--) {
    int gl = chain[i]; const int32_t* li = W->link_i + DG_LINK_I_W * gl; const double* lf = W->link_f + DG_LINK_F_W * gl;
    double E[9], r[3], t[3], Rn[9];
    joint_xform(W, gl, li[3] >= 0 ? qb[li[3] - d0] : 0.0, E, r);
    m_vec(t, Rc, r); v_add(pc, pc, t); m_mulT(Rn, Rc, E); memcpy(Rc, Rn, 72);
    jt[i] = li[2]; jd[i] = li[3] >= 0 ? li[3] - d0 : -1;
    m_vec(ax_w + 3 * i, Rc, lf + 10);
    double dw[3]; m_vec(dw, Rc, lf + 7); v_sub(org_w + 3 * i, pc, dw);
  }
  /* end-effector URDF link frame */
  const double* lfe = W->link_f + DG_LINK_F_W * ee_gl;
  double dw[3]; m_vec(dw, Rc, lfe + 7); v_sub(pos, pc, dw);
  double Rli[9], qi[4] = {-lfe[16], -lfe[17], -lfe[18], lfe[19]}; q_to_mat(Rli, qi); m_mul(R, Rc, Rli);
  for (int i = 0; i < 3 * ndb; i++) Jl[i] = Ja[i] = 0;
  for (int i = 0; i < nc; i++) {
    if (jd[i] < 0) continue;
    if (jt[i] == 1) { double rel[3], t[3]; v_sub(rel, pos, org_w + 3 * i); v_cross(t, ax_w + 3 * i, rel); for (int k = 0; k < 3; k++) { Jl[k * ndb + jd[i]] = t[k]; Ja[k * ndb + jd[i]] = ax_w[3 * i + k]; } }
    else if (jt[i] == 2) for (int k = 0; k < 3; k++) Jl[k * ndb + jd[i]] = ax_w[3 * i + k];
  }
}
static int solve_dense(double* A, double* bvec, int n) {  /* in-place Gaussian elimination with partial pivoting */
  for (int c = 0; c < n; c++) {
    int p = c; for (int r2 = c + 1; r2 < n; r2++) if (fabs(A[r2 * n + c]) > fabs(A[p * n + c])) p = r2;
    if (fabs(A[p * n + c]) < 1e-300) return -1;
    if (p != c) { for (int j = 0; j < n; j++) { double t = A[c * n + j]; A[c * n + j] = A[p * n + j]; A[p * n + j] = t; } double t = bvec[c]; bvec[c] = bvec[p]; bvec[p] = t; }
    for (int r2 = c + 1; r2 < n; r2++) { double f = A[r2 * n + c] / A[c * n + c]; if (f != 0) { for (int j = c; j < n; j++) A[r2 * n + j] -= f * A[c * n + j]; bvec[r2] -= f * bvec[c]; } }
    FL(n * n);
  }
  for (int r2 = n - 1; r2 >= 0; r2--) { double s = bvec[r2]; for (int j = r2 + 1; j < n; j++) s -= A[r2 * n + j] * bvec[j]; bvec[r2] = s / A[r2 * n + r2]; }
  return 0;
}
void dgo_ik(DgoWorld* W, int b, int ee_gl, const double* tpos_w, const double* torn_w, int use_orn, int nullspace,
            const double* lower, const double* upper, const double* range, const double* rest, double* out) {
  const int32_t* bi = W->body_i + DG_BODY_I_W * b; int d0 = bi[3], ndb = bi[4];
  double qb[64], Rb[9], tp[3], d[3], tq[4] = {0, 0, 0, 1};
  for (int i = 0; i < ndb; i++) qb[i] = ST(W, S_Q)[d0 + i];
  /* target in base coordinates */
  const double *bp = ST(W, S_BPOS) + 3 * b, *bq = ST(W, S_BQUAT) + 4 * b;
  q_to_mat(Rb, bq); v_sub(d, tpos_w, bp); mT_vec(tp, Rb, d);
  if (use_orn) { double bqi[4] = {-bq[0], -bq[1], -bq[2], bq[3]}; q_mul(tq, bqi, torn_w); }
  double nullv[64];
  if (nullspace) for (int i = 0; i < ndb; i++) {
    nullv[i] = 0.001 * (rest[i] - qb[i]);
    if (qb[i] > upper[i]) nullv[i] += 10.0 * (upper[i] - qb[i]) / range[i];
    if (qb[i] < lower[i]) nullv[i] += 10.0 * (lower[i] - qb[i]) / range[i];
  }
  int m = use_orn ? 6 : 3;
  double diff = 1e30, thr = HF(W, ik_threshold), lam = HF(W, ik_damping), lam2 = HF(W, ik_null_lambda_sq);
  for (int it = 0; it < W->ik_iters && diff > thr; it++) {
    double pos[3], R[9], Jl[3 * 64], Ja[3 * 64], J[6 * 64], e[6], dth[64];
    ik_fk(W, b, ee_gl, qb, pos, R, Jl, Ja);
    v_sub(e, tp, pos); diff = v_len(e);
    memcpy(J, Jl, sizeof(double) * 3 * (size_t)ndb);
    if (use_orn) {
      memcpy(J + 3 * ndb, Ja, sizeof(double) * 3 * (size_t)ndb);
      double qc[4], qci[4], dq[4]; mat_to_q(qc, R); qci[0] = -qc[0]; qci[1] = -qc[1]; qci[2] = -qc[2]; qci[3] = qc[3];
      q_mul(dq, tq, qci);
      /* dq is normalised before the acos: for unit inputs (what the reference hands to pybullet: fp64 quaternions from
         getLinkState x getQuaternionFromEuler) this changes nothing, but the parity protocol hands the oracle state rows
         ROUNDED TO FP32, whose quaternions are off unit norm by ~6e-8, and d(acos w)/dw = 1/sin(angle/2) ~ 300 for the
         0.4 deg rotations this controller asks for turns that into a 0.2 % error of the rotation vector (measured:
         7e-4 rad on the wrist targets after 20 iterations, 5e-3 rad/s on qdot - profiles/r2_qd_probe_*.json) */
      { double nq = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3]); if (nq > 0) for (int i = 0; i < 4; i++) dq[i] /= nq; }
      double wq = dq[3] > 1 ? 1 : (dq[3] < -1 ? -1 : dq[3]);
      double ang = 2 * acos(wq), s2 = 1 - wq * wq, ax[3];
      if (s2 < 10 * 2.220446049250313e-16) v_set(ax, 1, 0, 0); else v_scale(ax, dq, 1.0 / sqrt(s2));
      if (ang > M_PI) ang -= 2 * M_PI; else if (ang < -M_PI) ang += 2 * M_PI;
      double an = v_len(ax); if (an > 0) v_scale(ax, ax, 1.0 / an);
      v_scale(e + 3, ax, ang);
    }
    if (!nullspace) {            /* dth = (J^T J + lam I)^-1 J^T e */
      double A[64 * 64], rhs[64];
      for (int i = 0; i < ndb; i++) { for (int j = 0; j < ndb; j++) { double s = 0; for (int k = 0; k < m; k++) s += J[k * ndb + i] * J[k * ndb + j]; A[i * ndb + j] = s + (i == j ? lam : 0); } double s = 0; for (int k = 0; k < m; k++) s += J[k * ndb + i] * e[k]; rhs[i] = s; }
      FL(2 * m * ndb * ndb);
      if (solve_dense(A, rhs, ndb) != 0) break;
      memcpy(dth, rhs, sizeof(double) * (size_t)ndb);
    } else {                      /* dth = J^T (J J^T + lam2 I)^-1 e + (I - J^T (J J^T + lam2 I)^-1 J) nullv */
      double U[36], y[6], Jn[6];
      for (int i = 0; i < m; i++) { double s = 0; for (int k = 0; k < ndb; k++) s += J[i * ndb + k] * nullv[k]; Jn[i] = s; }
      for (int i = 0; i < m; i++) y[i] = e[i] - Jn[i];
      for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) { double s = 0; for (int k = 0; k < ndb; k++) s += J[i * ndb + k] * J[j * ndb + k]; U[i * m + j] = s + (i == j ? lam2 : 0); }
      FL(2 * m * m * ndb);
      if (solve_dense(U, y, m) != 0) break;
      for (int k = 0; k < ndb; k++) { double s = nullv[k]; for (int i = 0; i < m; i++) s += J[i * ndb + k] * y[i]; dth[k] = s; }
    }
    double mx = 0; for (int i = 0; i < ndb; i++) if (fabs(dth[i]) > mx) mx = fabs(dth[i]);
    double cap = 45.0 * M_PI / 180.0; if (mx > cap) for (int i = 0; i < ndb; i++) dth[i] *= cap / mx;
    for (int i = 0; i < ndb; i++) qb[i] += dth[i];
  }
  for (int i = 0; i < ndb; i++) out[i] = qb[i];
}

/* ---------------------------------------------------------------- add-on ops ---------------------------- */
static uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
static double urand(uint32_t seed, uint32_t env, uint32_t epoch, uint32_t stream) {
  uint32_t h = hash32(seed ^ hash32(env + 0x9e3779b9U * (epoch + 1)) ^ hash32(stream * 0x85ebca6bU + 0xc2b2ae35U));
  return (double)(h >> 8) * (1.0 / 16777216.0);
}
/* world-space axis and origin of the joint carrying link gl, from the cached link COM poses (getLinkState[0,1]) */
static void joint_axis_world(const DgoWorld* W, int gl, double* aw, double* ow) {
  const double* lf = W->link_f + DG_LINK_F_W * gl; double R[9], dw[3];
  q_to_mat(R, ST(W, S_LQUAT) + 4 * gl); m_vec(aw, R, lf + 10); m_vec(dw, R, lf + 7); v_sub(ow, ST(W, S_LPOS) + 3 * gl, dw);
}
/* diy_gym/addons/controllers/admittance_controller.py:36-55.  Restates the three pybullet calls it makes:
 *   p.calculateJacobian(uid, end_frame, offset, q, 0, 0)  -> translational / rotational Jacobian of the point `offset`
 *       (given in the link's COM frame) in world axes, one column per DoF  [RECALLED-UNVERIFIED frame conventions];
 *   p.calculateInverseDynamics(uid, q, 0, 0)              -> G(q), the generalized gravity forces of a fixed-base body;
 *   p.setJointMotorControlArray(..., TORQUE_CONTROL, forces = F.J_lin + T.J_ang + G + kp (target - q) - kd qd). */
static void admittance_update(DgoWorld* W, const int32_t* ia, const double* fa, const double* a) {
  int b = ia[0], ee = ia[1], n = ia[2]; const int32_t* bi = W->body_i + DG_BODY_I_W * b; int l0 = bi[1], nlb = bi[2];
  double kp = fa[0], kd = fa[1]; const double* target = fa + 5;
  double g[3] = {HF(W, gx), HF(W, gy), HF(W, gz)};
  for (int i = 0; i < n; i++) { int d = ia[3 + i]; ST(W, S_JTORQUE)[d] += (target[i] - ST(W, S_Q)[d]) * kp - kd * ST(W, S_QD)[d]; }
  double Re[9], P[3], t[3];
  q_to_mat(Re, ST(W, S_LQUAT) + 4 * ee); m_vec(t, Re, fa + 2); v_add(P, ST(W, S_LPOS) + 3 * ee, t);
  for (int gl = ee; gl >= 0; gl = W->link_i[DG_LINK_I_W * gl + 1]) {
    const int32_t* li = W->link_i + DG_LINK_I_W * gl;
    if (li[3] < 0) continue;
    double aw[3], ow[3], rel[3], c[3];
    joint_axis_world(W, gl, aw, ow);
    if (li[2] == 1) { v_sub(rel, P, ow); v_cross(c, aw, rel); ST(W, S_JTORQUE)[li[3]] += v_dot(a, c) + v_dot(a + 3, aw); }
    else ST(W, S_JTORQUE)[li[3]] += v_dot(a, aw);
  }
  for (int k = 0; k < nlb; k++) {
    double m = PR(W, P_MASS)[W->nb + l0 + k];
    if (m == 0) continue;
    const double* pk = ST(W, S_LPOS) + 3 * (l0 + k);
    for (int gl = l0 + k; gl >= 0; gl = W->link_i[DG_LINK_I_W * gl + 1]) {
      const int32_t* li = W->link_i + DG_LINK_I_W * gl;
      if (li[3] < 0) continue;
      double aw[3], ow[3], rel[3], c[3];
      joint_axis_world(W, gl, aw, ow);
      if (li[2] == 1) { v_sub(rel, pk, ow); v_cross(c, aw, rel); } else v_cpy(c, aw);
      ST(W, S_JTORQUE)[li[3]] -= m * v_dot(g, c);
    }
  }
}
/* controllers: diy_gym/addons/controllers/ update() bodies */
void dgo_apply_actions(DgoWorld* W, const double* act) {
  for (int k = 0; k < W->nop; k++) {
    const int32_t* op = W->op_i + DG_OP_I_W * k; const int32_t* ia = W->oparg_i + op[1]; const double* fa = W->oparg_f + op[2];
    const double* a = op[3] >= 0 ? act + op[3] : NULL;
    if (op[0] == OP_JOINT_CTRL) {                       /* joint_controller.py:40-58 */
      int mode = ia[0], n = ia[1];
      for (int i = 0; i < n; i++) {
        int d = ia[2 + i];
        if (mode == 2) { ST(W, S_JTORQUE)[d] += a[i]; continue; }
        ST(W, S_MKD)[d] = fa[1]; ST(W, S_MMAXF)[d] = fa[2 + i];
        if (mode == 0) { ST(W, S_MKP)[d] = fa[0]; ST(W, S_MTPOS)[d] = a[i]; ST(W, S_MTVEL)[d] = 0; }
        else { ST(W, S_MKP)[d] = 0; ST(W, S_MTPOS)[d] = 0; ST(W, S_MTVEL)[d] = a[i]; }
      }
    } else if (op[0] == OP_EXT_FORCE) {                 /* external_force.py:21-24 */
      int f = ia[0]; double pos[3], quat[4], v[3], o[3], F[3], rel[3], t[3];
      frame_com_state(W, f, pos, quat, v, o);
      if (ia[1] == 0) { v_cpy(F, a); v_sub(rel, fa, pos); }
      else { double R[9]; q_to_mat(R, quat); m_vec(F, R, a); m_vec(rel, R, fa); }
      v_add(ST(W, S_EXTF) + 3 * f, ST(W, S_EXTF) + 3 * f, F); v_cross(t, rel, F); v_add(ST(W, S_EXTT) + 3 * f, ST(W, S_EXTT) + 3 * f, t);
    } else if (op[0] == OP_FILTERED_WRENCH) {           /* examples/drone_pilot/drone_pilot.py:31-37 (Propellor.update): first-order
                                                           filter of the action, applyExternalForce / Torque in LINK_FRAME scaled by it */
      double* st_ = ST(W, S_ADDON) + ia[2]; st_[0] = st_[0] + (a[0] - st_[0]) * fa[0];
      int f = ia[0]; double pos[3], quat[4], v[3], o[3], R[9], F[3], T[3], rel[3], t[3];
      frame_com_state(W, f, pos, quat, v, o); q_to_mat(R, quat);
      double Fl[3] = {fa[1] * st_[0], fa[2] * st_[0], fa[3] * st_[0]}, Tl[3] = {fa[4] * st_[0], fa[5] * st_[0], fa[6] * st_[0]};
      m_vec(F, R, Fl); m_vec(T, R, Tl); m_vec(rel, R, fa + 7);
      v_add(ST(W, S_EXTF) + 3 * f, ST(W, S_EXTF) + 3 * f, F); v_cross(t, rel, F); v_add(t, t, T); v_add(ST(W, S_EXTT) + 3 * f, ST(W, S_EXTT) + 3 * f, t);
    } else if (op[0] == OP_ADMITTANCE) {                /* admittance_controller.py:36-55 */
      admittance_update(W, ia, fa, a);
    } else if (op[0] == OP_IK_CTRL) {                   /* ik_controller.py:51-80 */
      int b = ia[0], ee = ia[1], n = ia[2], use_orn = ia[3], ns = ia[4]; int ndb = W->body_i[DG_BODY_I_W * b + 4];
      double pos[3], quat[4], v[3], o[3], tq[4], out[64];
      frame_com_state(W, W->nb + ee, pos, quat, v, o);
      /* target = current link COM pose (getLinkState[0],[1]) + delta, handed to the IK as a link-frame target exactly as
         the reference does (ik_controller.py:52-59); the end-effector links of the example robots have no inertial offset */
      double tpos[3] = {pos[0] + a[0], pos[1] + a[1], pos[2] + a[2]};
      if (use_orn) { double dq[4]; q_from_euler(dq, a + 3); q_mul(tq, quat, dq); }
      const double* lim = fa + 2 + n;
      dgo_ik(W, b, ee, tpos, tq, use_orn, ns, lim, lim + ndb, lim + 2 * ndb, lim + 3 * ndb, out);
      for (int i = 0; i < n; i++) {
        int d = ia[5 + i];
        ST(W, S_MKP)[d] = fa[0]; ST(W, S_MKD)[d] = fa[1]; ST(W, S_MMAXF)[d] = fa[2 + i]; ST(W, S_MTPOS)[d] = out[i]; ST(W, S_MTVEL)[d] = 0;
      }
    }
  }
}
/* sensors / rewards / terminals: diy_gym/addons/{sensors,rewards}/ */
void dgo_observe(DgoWorld* W, double* obs, double* rew, uint8_t* term) {
  double dt = HF(W, dt);
  for (int k = 0; k < W->nop; k++) {
    const int32_t* op = W->op_i + DG_OP_I_W * k; const int32_t* ia = W->oparg_i + op[1]; const double* fa = W->oparg_f + op[2];
    double* o = op[4] >= 0 ? obs + op[4] : NULL;
    if (op[0] == OP_JOINT_SENSOR) {                     /* joint_state_sensor.py:46-57 */
      int n = ia[0], flags = ia[1], j = 0;
      for (int i = 0; i < n; i++) o[j++] = ST(W, S_Q)[ia[2 + i]];
      if (flags & 1) for (int i = 0; i < n; i++) o[j++] = ST(W, S_QD)[ia[2 + i]];
      if (flags & 2) for (int i = 0; i < n; i++) o[j++] = ST(W, S_MAPPLIED)[ia[2 + i]] / dt;
    } else if (op[0] == OP_OBJECT_SENSOR) {             /* object_state_sensor.py:33-75 */
      double p[3], q[4], v[3], w[3]; int flags = ia[2], j = 0;
      frame_com_state(W, ia[0], p, q, v, w);
      if (ia[1] >= 0) {
        double sp[3], sq[4], sv[3], sw[3], qq[4];
        frame_com_state(W, ia[1], sp, sq, sv, sw);
        v_sub(p, p, sp); v_sub(v, v, sv); q_mul(qq, sq, q); memcpy(q, qq, 32); v_sub(w, w, sw);
      }
      for (int i = 0; i < 3; i++) o[j++] = p[i];
      if (flags & 2) for (int i = 0; i < 3; i++) o[j++] = v[i];
      if (flags & 1) { double e[3]; euler_from_q(e, q); for (int i = 0; i < 3; i++) o[j++] = e[i]; }
      if ((flags & 3) == 3) for (int i = 0; i < 3; i++) o[j++] = w[i];
    } else if (op[0] == OP_FT_SENSOR) {                 /* force_torque_sensor.py:21-23 */
      for (int i = 0; i < 6; i++) o[i] = ST(W, S_JREACT)[6 * ia[0] + i];
    } else if (op[0] == OP_REACH_TARGET) {              /* reach_target.py:21-36 */
      double sp[3], sq[4], tp[3], tq[4], d[3];
      frame_link_pose(W, ia[0], sp, sq); frame_link_pose(W, ia[1], tp, tq); v_sub(d, tp, sp);
      double dist = v_len(d);
      rew[op[5]] = -dist * fa[0]; term[op[6]] = dist < fa[1];
    } else if (op[0] == OP_ELECTRICITY) {               /* electricity_cost.py:15-18 */
      const int32_t* bi = W->body_i + DG_BODY_I_W * ia[0]; double s = 0;
      for (int i = 0; i < bi[4]; i++) s += fabs(ST(W, S_MAPPLIED)[bi[3] + i] / dt * ST(W, S_QD)[bi[3] + i]);
      rew[op[5]] = -s * fa[0];
    } else if (op[0] == OP_STUCK_JOINT) {               /* stuck_joint_cost.py:19-21 (intent; the reference raises NameError) */
      const int32_t* bi = W->body_i + DG_BODY_I_W * ia[0]; int stuck = 0;
      for (int l = 0; l < bi[2]; l++) {
        const int32_t* li = W->link_i + DG_LINK_I_W * (bi[1] + l); const double* lf = W->link_f + DG_LINK_F_W * (bi[1] + l);
        if (li[3] < 0) continue;
        double qq = ST(W, S_Q)[li[3]];
        if (fmin(fabs(lf[20] - qq), fabs(lf[21] - qq)) < 0.01) stuck = 1;
      }
      rew[op[5]] = stuck ? -fa[0] : 0.0;
    } else if (op[0] == OP_TIME_PENALTY) {              /* time_penalty.py:11-12 */
      rew[op[5]] = fa[0];
    } else if (op[0] == OP_FILTERED_WRENCH) {           /* Propellor.observe (drone_pilot.py:39-40) */
      o[0] = ST(W, S_ADDON)[ia[2]];
    } else if (op[0] == OP_TILT_TERMINAL) {             /* FellOver.is_terminal (drone_pilot.py:52-55) */
      const double* q = ST(W, S_BQUAT) + 4 * ia[0];
      term[op[6]] = 2.0 * atan2(sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]), fabs(q[3])) > fa[0];
    } else if (op[0] == OP_EPISODE_TIMER) {             /* diy_gym.py:180-183 */
      term[op[6]] = ST(W, S_STEP)[0] >= fa[0];
    }
  }
}
/* reset(): diy_gym/diy_gym.py:130-148 with the add-on reset() bodies of joint/ik controllers, respawn, dynamics_randomizer */
void dgo_env_reset(DgoWorld* W) {
  ST(W, S_STEP)[0] = 0;
  uint32_t epoch = (uint32_t)ST(W, S_RESETS)[0];
  for (int k = 0; k < W->nop; k++) {
    const int32_t* op = W->op_i + DG_OP_I_W * k; const int32_t* ia = W->oparg_i + op[1]; const double* fa = W->oparg_f + op[2];
    if (op[0] == OP_FILTERED_WRENCH) { if (ia[1]) ST(W, S_ADDON)[ia[2]] = 0; }
    else if (op[0] == OP_JOINT_RESET) {                 /* joint_controller.py:36-38, ik_controller.py:47-49 */
      for (int i = 0; i < ia[0]; i++) { ST(W, S_Q)[ia[1 + i]] = fa[i]; ST(W, S_QD)[ia[1 + i]] = 0; }
    } else if (op[0] == OP_RESPAWN) {                   /* respawn.py:31-39 */
      int b = ia[0]; uint32_t ep = ia[1] ? 0u : epoch; const double* ip = PR(W, P_INITPOSE) + 7 * b;
      double e[3], dq[4], qo[4];
      for (int i = 0; i < 3; i++) ST(W, S_BPOS)[3 * b + i] = ip[i] + (urand(W->seed, (uint32_t)W->env_id, ep, (uint32_t)(k * 8 + i)) - 0.5) * fa[i];
      for (int i = 0; i < 3; i++) e[i] = (urand(W->seed, (uint32_t)W->env_id, ep, (uint32_t)(k * 8 + 3 + i)) - 0.5) * fa[3 + i];
      q_from_euler(dq, e); q_mul(qo, ip + 3, dq); memcpy(ST(W, S_BQUAT) + 4 * b, qo, 32);
      v_set(ST(W, S_BVEL) + 3 * b, 0, 0, 0); v_set(ST(W, S_BOMEGA) + 3 * b, 0, 0, 0);
    } else if (op[0] == OP_VIS_RANDOMIZE) {             /* visual_randomizer.py:41-46: a new look per reset - a random colour per visual shape of the
                                                           body (the reference's random texture comes from a dataset it downloads) */
      for (int v = 0; v < W->nv; v++) if (W->vis_i[DG_VIS_I_W * v + 3] == ia[0])
        for (int i = 0; i < 3; i++) PR(W, P_COLOR)[3 * v + i] = urand(W->seed, (uint32_t)W->env_id, epoch, (uint32_t)(k * 8 + 128 + 3 * v + i));
    } else if (op[0] == OP_DYN_RANDOMIZE) {             /* dynamics_randomizer.py:24-32 (log-uniform on nominal values) */
      int b = ia[0]; const int32_t* bi = W->body_i + DG_BODY_I_W * b;
      for (int l = -1; l < bi[2]; l++) {
        int f = l < 0 ? b : W->nb + bi[1] + l; int d = l < 0 ? -1 : W->link_i[DG_LINK_I_W * (bi[1] + l) + 3];
        if (l >= 0 && d < 0) continue;
        if (l < 0 && bi[4] > 0) continue;
        double u1 = urand(W->seed, (uint32_t)W->env_id, epoch, (uint32_t)(k * 8 + 64 + 2 * (l + 1)));
        double u2 = urand(W->seed, (uint32_t)W->env_id, epoch, (uint32_t)(k * 8 + 65 + 2 * (l + 1)));
        double ms = exp(log(fa[0]) + u1 * (log(fa[1]) - log(fa[0]))), ds = exp(log(fa[2]) + u2 * (log(fa[3]) - log(fa[2])));
        PR(W, P_MASS)[f] = W->param_def[HI(W, P_MASS) + f] * ms;
        for (int i = 0; i < 3; i++) PR(W, P_INERTIA)[3 * f + i] = W->param_def[HI(W, P_INERTIA) + 3 * f + i] * ms;
        /* fa[6]: nominal joint damping for joints whose URDF gives none (UR5: 0 - the scaling would be a no-op; extension key) */
        if (d >= 0) { double nom = W->param_def[HI(W, P_JDAMP) + d]; PR(W, P_JDAMP)[d] = (nom > 0 ? nom : fa[6]) * ds; }
      }
      /* extension key friction_range (BASELINE.json config 5, SURVEY 8d: lateral friction U[0.5, 1.25] per environment): one draw per
         body and reset, every collision shape of the body gets it */
      if (fa[5] > 0) {
        double u3 = urand(W->seed, (uint32_t)W->env_id, epoch, (uint32_t)(k * 8 + 63));
        double fr = fa[4] + u3 * (fa[5] - fa[4]);
        for (int s2 = 0; s2 < W->ns; s2++) if (W->shape_i[DG_SHAPE_I_W * s2] == b) PR(W, P_FRICTION)[s2] = fr;
      }
    }
  }
  ST(W, S_RESETS)[0] += 1;
  write_link_cache(W);
  for (int i = 0; i < W->hot_start; i++) dgo_step_physics(W);
}
/* step(): diy_gym/diy_gym.py:187-209 */
void dgo_env_step(DgoWorld* W, const double* act, double* obs, double* rew, uint8_t* term) {
  dgo_apply_actions(W, act);
  ST(W, S_STEP)[0] += 1;
  dgo_step_physics(W);
  dgo_observe(W, obs, rew, term);
}

/* ---------------------------------------------------------------- camera (ray cast) --------------------- */
static int ray_shape(int type, const double* d, const double* o, const double* dir, double tmax, double* t_out, double* n_out) {
  /* ray in shape-local coordinates; returns nearest hit t in (1e-9, tmax) */
  double best = tmax; int hit = 0; double nb_[3] = {0, 0, 1};
  if (type == SHAPE_SPHERE || type == SHAPE_CAPSULE) {
    int nsph = type == SHAPE_SPHERE ? 1 : 2;
    for (int s = 0; s < nsph; s++) {
      double c[3] = {0, 0, type == SHAPE_SPHERE ? 0 : (s ? d[1] : -d[1])}, oc[3]; v_sub(oc, o, c);
      double A = v_dot(dir, dir), B = v_dot(oc, dir), C = v_dot(oc, oc) - d[0] * d[0], disc = B * B - A * C;
      if (disc < 0) continue;
      double t = (-B - sqrt(disc)) / A;
      if (t > 1e-9 && t < best) { best = t; hit = 1; for (int i = 0; i < 3; i++) nb_[i] = (oc[i] + t * dir[i]) / d[0]; }
    }
  }
  if (type == SHAPE_CAPSULE || type == SHAPE_CYLINDER) {
    double A = dir[0] * dir[0] + dir[1] * dir[1], B = o[0] * dir[0] + o[1] * dir[1], C = o[0] * o[0] + o[1] * o[1] - d[0] * d[0];
    double disc = B * B - A * C;
    if (A > 1e-18 && disc >= 0) {
      double t = (-B - sqrt(disc)) / A, z = o[2] + t * dir[2];
      if (t > 1e-9 && t < best && fabs(z) <= d[1]) { best = t; hit = 1; nb_[0] = (o[0] + t * dir[0]) / d[0]; nb_[1] = (o[1] + t * dir[1]) / d[0]; nb_[2] = 0; }
    }
    if (type == SHAPE_CYLINDER && fabs(dir[2]) > 1e-18) for (int s = -1; s <= 1; s += 2) {
      double t = (s * d[1] - o[2]) / dir[2], x = o[0] + t * dir[0], y = o[1] + t * dir[1];
      if (t > 1e-9 && t < best && x * x + y * y <= d[0] * d[0]) { best = t; hit = 1; nb_[0] = 0; nb_[1] = 0; nb_[2] = s; }
    }
  }
  if (type == SHAPE_BOX) {
    double t0 = -1e300, t1 = 1e300; int ax0 = 0; double sg0 = 1;
    for (int i = 0; i < 3; i++) {
      if (fabs(dir[i]) < 1e-18) { if (fabs(o[i]) > d[i]) return 0; continue; }
      double ta = (-d[i] - o[i]) / dir[i], tb = (d[i] - o[i]) / dir[i], sg = -1;
      if (ta > tb) { double t = ta; ta = tb; tb = t; sg = 1; }
      if (ta > t0) { t0 = ta; ax0 = i; sg0 = sg; }
      if (tb < t1) t1 = tb;
    }
    if (t0 <= t1 && t0 > 1e-9 && t0 < best) { best = t0; hit = 1; nb_[0] = nb_[1] = nb_[2] = 0; nb_[ax0] = sg0; }
  }
  if (hit) { *t_out = best; v_cpy(n_out, nb_); }
  return hit;
}
/* camera.py:58-92: rgb (float in [0,1]) and eye-space depth (negative z, as the reference's linearisation yields).
 * Buffers are row-major image rows (height x width), which the reference then labels (W,H,...) (camera.py:77,82). */
void dgo_render_seg(DgoWorld* W, int cam, double* rgb, double* depth, double* seg) {
  const int32_t* ci = W->cam_i + DG_CAM_I_W * cam; const double* cf = W->cam_f + DG_CAM_F_W * cam;
  int width = ci[1], height = ci[2];
  double Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pp[3] = {0, 0, 0};
  if (ci[0] >= 0) { double q[4]; frame_link_pose(W, ci[0], pp, q); q_to_mat(Rp, q); }
  double Rl[9], Rc[9], pc[3], t[3];
  q_to_mat(Rl, cf + 3); m_mul(Rc, Rp, Rl); m_vec(t, Rp, cf); v_add(pc, pp, t);
  double fov = cf[7], nearp = cf[8], farp = cf[9], th = tan(fov * M_PI / 360.0), aspect = (double)width / height;
  /* visual shapes in world */
  int nv = W->nv; double* VR = (double*)malloc(sizeof(double) * 12 * (size_t)(nv > 0 ? nv : 1));
  for (int s = 0; s < nv; s++) {
    const int32_t* vi = W->vis_i + DG_VIS_I_W * s; const double* vf = W->vis_f + DG_VIS_F_W * s;
    double p[3], q[4], v[3], o[3], R[9], Rs[9];
    frame_com_state(W, vi[0], p, q, v, o); q_to_mat(R, q); q_to_mat(Rs, vf + 3); m_mul(VR + 12 * s, R, Rs);
    m_vec(t, R, vf); v_add(VR + 12 * s + 9, p, t);
  }
  const double light[3] = {0.4082482904638631, 0.4082482904638631, 0.8164965809277261};
  for (int j = 0; j < height; j++) for (int i = 0; i < width; i++) {
    double dc[3] = {((i + 0.5) / width * 2 - 1) * th * aspect, (1 - (j + 0.5) / height * 2) * th, -1.0}, dw[3];
    m_vec(dw, Rc, dc);
    double best = farp; int hs = -1; double hn[3] = {0, 0, 1};
    for (int s = 0; s < nv; s++) {
      const int32_t* vi = W->vis_i + DG_VIS_I_W * s; const double* vf = W->vis_f + DG_VIS_F_W * s;
      double oc[3], ol[3], dl[3], tt, nn[3];
      v_sub(oc, pc, VR + 12 * s + 9); mT_vec(ol, VR + 12 * s, oc); mT_vec(dl, VR + 12 * s, dw);
      if (ray_shape(vi[1], vf + 7, ol, dl, best, &tt, nn) && tt >= nearp) { best = tt; hs = s; m_vec(hn, VR + 12 * s, nn); }
    }
    int px = j * width + i;
    if (seg) seg[px] = hs < 0 ? -1.0 : (double)W->vis_i[DG_VIS_I_W * hs + 3];   /* camera.py:89-90: unique id of the visible body */
    if (hs < 0) { rgb[3 * px] = rgb[3 * px + 1] = rgb[3 * px + 2] = 1.0; depth[px] = -farp; }
    else {
      const double* col = PR(W, P_COLOR) + 3 * hs; double nl = v_dot(hn, light); if (nl < 0) nl = 0;   /* per-environment colour (visual_randomizer) */
      double sh = 0.4 + 0.6 * nl;
      for (int k = 0; k < 3; k++) rgb[3 * px + k] = col[k] * sh;
      depth[px] = -best;
    }
  }
  free(VR);
}
void dgo_render(DgoWorld* W, int cam, double* rgb, double* depth) { dgo_render_seg(W, cam, rgb, depth, (double*)0); }

/* p.getCameraImage(width, height, viewMatrix, projectionMatrix) as the reference calls it (camera.py:70-74): the camera is given
 * ONLY by the two column-major 4x4 OpenGL matrices the caller hands over - nothing of this repo's camera tables is consulted, so a
 * pose / field-of-view / row-order convention that differs between the reference's add-on and the compiled camera shows up as a
 * different image.  Returns what pybullet returns: rgba bytes [height][width][4], the NON-linear depth buffer in [0,1]
 * (z_ndc * 0.5 + 0.5, which camera.py:80-85 turns back into eye-space z) and the body id per pixel (-1 background).
 * Shading and intersection routines are those of dgo_render_seg (the renderer itself is this repo's, TinyRenderer is absent). */
void dgo_get_camera_image(DgoWorld* W, int width, int height, const double* view, const double* proj, unsigned char* rgba, double* depth01, int* segm) {
  /* view = [R^T | -R^T p] column-major: eye axes are the rows of its rotation block */
  double Rc[9], pc[3], tv[3] = {view[12], view[13], view[14]};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rc[3 * i + j] = view[4 * i + j];     /* world_from_eye rotation = (R^T)^T */
  for (int i = 0; i < 3; i++) pc[i] = -(Rc[3 * i] * tv[0] + Rc[3 * i + 1] * tv[1] + Rc[3 * i + 2] * tv[2]);
  const double p00 = proj[0], p11 = proj[5], p22 = proj[10], p32 = proj[14];                   /* column-major: proj[4 c + r] */
  const double nearp = p32 / (p22 - 1.0), farp = p32 / (p22 + 1.0);
  int nv = W->nv; double* VR = (double*)malloc(sizeof(double) * 12 * (size_t)(nv > 0 ? nv : 1)); double t[3];
  for (int s = 0; s < nv; s++) {
    const int32_t* vi = W->vis_i + DG_VIS_I_W * s; const double* vf = W->vis_f + DG_VIS_F_W * s;
    double p[3], q[4], v[3], o[3], R[9], Rs[9];
    frame_com_state(W, vi[0], p, q, v, o); q_to_mat(R, q); q_to_mat(Rs, vf + 3); m_mul(VR + 12 * s, R, Rs);
    m_vec(t, R, vf); v_add(VR + 12 * s + 9, p, t);
  }
  const double light[3] = {0.4082482904638631, 0.4082482904638631, 0.8164965809277261};
  for (int j = 0; j < height; j++) for (int i = 0; i < width; i++) {
    /* pixel centre -> normalised device coordinates (row 0 is the top of the image) -> eye-space direction with z = -1 */
    const double xn = (i + 0.5) / width * 2 - 1, yn = 1 - (j + 0.5) / height * 2;
    double dc[3] = {xn / p00, yn / p11, -1.0}, dw[3];
    m_vec(dw, Rc, dc);
    double best = farp; int hs = -1; double hn[3] = {0, 0, 1};
    for (int s = 0; s < nv; s++) {
      const int32_t* vi = W->vis_i + DG_VIS_I_W * s; const double* vf = W->vis_f + DG_VIS_F_W * s;
      double oc[3], ol[3], dl[3], tt, nn[3];
      v_sub(oc, pc, VR + 12 * s + 9); mT_vec(ol, VR + 12 * s, oc); mT_vec(dl, VR + 12 * s, dw);
      if (ray_shape(vi[1], vf + 7, ol, dl, best, &tt, nn) && tt >= nearp) { best = tt; hs = s; m_vec(hn, VR + 12 * s, nn); }
    }
    const int px = j * width + i;
    double col[3] = {1, 1, 1};
    if (hs >= 0) { const double* c = PR(W, P_COLOR) + 3 * hs; double nl = v_dot(hn, light); if (nl < 0) nl = 0; const double sh = 0.4 + 0.6 * nl; for (int k = 0; k < 3; k++) col[k] = c[k] * sh; }
    for (int k = 0; k < 3; k++) { double b = floor(col[k] * 255.0 + 0.5); rgba[4 * px + k] = (unsigned char)(b < 0 ? 0 : (b > 255 ? 255 : b)); }
    rgba[4 * px + 3] = 255;
    /* eye depth `best` -> z_ndc = (p22 * (-best) + p32) / best -> depth buffer value */
    depth01[px] = 0.5 * ((p22 * (-best) + p32) / best) + 0.5;
    if (segm) segm[px] = hs < 0 ? -1 : W->vis_i[DG_VIS_I_W * hs + 3];
  }
  free(VR);
}
