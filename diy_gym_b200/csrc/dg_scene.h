// dg_scene.h - device-visible scene description and the per-environment workspace plan.
//
// The Python compiler (diy_gym_b200/compiler/scene.py) hands the C ABI two flat buffers (int32 + float64
// sections, layout in scene_sections.h).  `HostScene::build` converts them to fp32, derives the tables the
// kernels need (dynamic-body list, frame slots, depth of every link, baked world poses of static shapes)
// and lays out the per-environment workspace that the step kernel keeps in shared memory.
//
// This replaces what the reference keeps inside its physics server after `p.loadURDF`
// (/root/reference/diy_gym/model.py:65) - here it is plain constant arrays shared by all environments.
#pragma once
#include <stdint.h>

#include "scene_sections.h"

namespace dg {

// per-dof hot arrays in the workspace: q, qd, applied motor impulse, joint torque (applied + damping, fixed for the step)
enum { D_Q = 0, D_QD, D_APPLIED, D_TAU, D_COUNT };
// body plan columns
enum { BP_DI = 0, BP_GDIM, BP_GVOFF, BP_MINVOFF, BP_I0OFF, BP_SLOT, BP_UROW, BP_DEPTH, BP_NRS, BP_AOFF, BP_GS, BP_W };
// workspace header ints
enum { WH_NCONTACT = 0, WH_NCROW, WH_NSURV, WH_COUPLED, WH_RS_R, WH_RS_NU, WH_RS_K, WH_RS_NEED, WH_RS_DEFER, WH_SUB, WH_COUNT = 12 };
// stages of one DIYGym.step when the contact sweeps run in their own kernel (dg_kernels.cu): see dg_env.cuh "the phase schedule"
enum { ST_ACT = 1, ST_PRE = 2, ST_POST = 4, ST_END = 8, ST_LOADC = 16, ST_SAVEC = 32, ST_LOAD = 128, ST_ALL = 15 };
// classes of the sweep kernel (dg_kernels.cu, dg_solve_kernel<W, K>): W lanes per environment, K consecutive rows per lane, W K row
// positions; an environment goes to the first class its K-padded layout fits (DevScene::rs_cls_k / rs_cls_r)
enum { RS_NCLS = 3 };
// thread-local scratch of minv_column: links per body, tree depth + 2
enum { DG_MINV_MAXL = 64, DG_MINV_MAXD = 24 };
// row table of the row-space team solver: contact row rr = RS_CONTACT | rr, unit row j of dynamic body di = di << 16 | j
enum { RS_CONTACT = 0x40000000, RS_KMAX = 8, RS_GVMAX = 128 };
// row record of the row-space solver
enum { RR_RHS = 0, RR_DINV, RR_LO, RR_HI, RR_MU, RR_PAR, RR_APPLIED, RR_ID, RR_W = 8 };
// row record strides
enum { UR_RHS = 0, UR_DINV, UR_LO, UR_HI, UR_APPLIED, UR_COL, UR_MOTOR, UR_W = 8 };
enum { CR_RHS = 0, CR_DINV, CR_LO, CR_HI, CR_APPLIED, CR_MU, CR_DA, CR_DB, CR_PARENT, CR_HDR = 12 };
// per-slot record strides of the cold workspace (multiples of 4 floats, so that records load with 128-bit accesses)
enum { LK_W = 20, AB_PA = 0, AB_WV = 6, AB_A = 12, AB_B = 21, AB_C = 30, AB_ACC = 40, AB_W = 48, LX_W = 8 };
enum { CT_FA = 0, CT_FB, CT_PA = 2, CT_PB = 5, CT_N = 8, CT_DIST = 11, CT_MU = 12, CT_W = 13 };

struct DevScene {
  int nb, nl, nd, ns, nv, npair, ncam, nop, nframes, S, P, substeps, iters, maxc, hot_start, ik_iters;
  int n_act, n_obs, n_rew, n_term;
  int so[24];  // state offsets:  so[HI_S_x - HI_S_BPOS]
  int po[12];  // param offsets:  po[HI_P_x - HI_P_MASS]
  float dt, g[3], erp, cerp, slop, margin, ik_damping, ik_threshold, max_joint_vel, limit_max_impulse, ik_null_lambda_sq;
  const int *body_i, *link_i, *shape_i, *pair_i, *vis_i, *op_i, *oparg_i, *cam_i;
  const float *body_f, *link_f, *shape_f, *vis_f, *oparg_f, *cam_f, *param_def, *state_def;
  const float* hull_f;    // reduced convex hulls of mesh collision shapes (shape_i[4..7]: vertex offset, vertices, plane offset, planes)
  // ---- plan ----
  int ndyn, nslot, nshw, nfloat, GD, GP, max_depth, max_nlb, n_ik, team;
  const int* dyn_body;    // [ndyn] body index of every body with kind != 0
  const int* body_plan;   // [nb][BP_W]
  const int* frame_slot;  // [nframes] workspace slot of a frame, -1 for frames of baked static bodies
  const int* link_depth;  // [nl] depth below the base (children of the base = 0)
  const int* shape_slot;  // [ns] index into the per-env shape-world array, -1 when baked
  const float* shape_wb;  // [ns][12] world rotation (9) + centre (3) of baked shapes
  const float* vis_wb;    // [nv][12] same for baked visual shapes
  const float* link_x;    // [nl][16] joint rest rotation R0 (9), motion subspace angular (3) / linear (3), pad
  // broad-phase groups: all candidate pairs between one baked static shape and one dynamic body share a
  // (static shape) x (body bounding sphere) pre-test; the remaining pairs are tested one by one
  int ngrp, nloose;
  const int* grp_i;       // [ngrp][4] static shape, dynamic body, first entry in grp_pairs, count
  const int* grp_pairs;   // pair indices, grouped
  const int* loose_pairs; // [nloose] pair indices outside any group
  const float* body_reach;// [nb] radius about the base COM that contains every collision shape of the body
  // ---- workspace layout (float offsets) ----
  int w_total, g_total, ws_mode;   // hot (shared) and cold (global) floats per team
  int W_HDR, W_BST, W_DOF, W_EXT, W_KIN, W_LINK, W_MINV, W_DV, W_I0, W_UCNT, W_CAPP, W_X;
  // region X, articulated-body phase
  int X_ABA, X_LNK, X_I0T;
  // region X, constraint phase
  int X_SHW, X_CON, X_SURV, X_CTMP, X_UROW, X_CROW, X_MSCR, X_AMAT, X_IK, X_RSA, X_RSAS, X_RSV, X_RSREC;
  int GV;           // generalized coordinates of all dynamic bodies, each padded to a multiple of 4 (the layout of W_DV)
  int rs_ashared;   // floats of shared memory per environment for the row-space matrix A (environments whose A is larger keep it in the cold workspace)
  int rs_cap;   // row capacity of the row-space team solver (0: unavailable, team of one lane)
  int ncons;        // fixed constraints between bodies (6 solver rows each, row-space solver only)
  const int* cons_i; const float* cons_f;
  int need_react;   // a force / torque sensor op reads the joint reaction wrenches (state section S_JREACT)
  int rs_min;   // contact rows an uncoupled environment needs before the team solves it in row space (fewer: per-body sweeps)
  int rs_cls_k[RS_NCLS], rs_cls_r[RS_NCLS];   // sweep-kernel classes: rows per lane, row positions (0: class unused)
  int solver;   // 1: contact environments are solved in row space by the whole team (default), 0: per-body dv-space sweeps
  int crow_stride, mscr_stride, ctmp_stride, ik_stride;
  int sem;      // SEM_* switches of the engine semantics that could only be recalled (compiler/scene.py SEMANTICS)
  int precise;  // 1: libm sincosf in FK / IK instead of the SFU approximation (scenes with fixed constraints; DG_PRECISE=0/1 overrides)
};

#define DG_SO(sc, name) ((sc)->so[HI_##name - HI_S_BPOS])
#define DG_PO(sc, name) ((sc)->po[HI_##name - HI_P_MASS])

}  // namespace dg

#ifndef __CUDACC_RTC__
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

namespace dg {

// Host-side owner of the converted scene.  `dev` holds HOST pointers into the vectors below; the CUDA
// backend re-targets them to device copies, the test-only emulation uses them as they are.
struct HostScene {
  std::vector<int> ints;      // all int tables, concatenated
  std::vector<float> floats;  // all float tables, concatenated
  DevScene dev;
  std::string error;

  // offsets of every table inside ints / floats (so a device copy can be re-pointed)
  struct Off { size_t body_i, link_i, shape_i, pair_i, vis_i, op_i, oparg_i, cam_i, dyn_body, body_plan, frame_slot, link_depth, shape_slot, grp_i, grp_pairs, loose_pairs;
               size_t body_f, link_f, shape_f, vis_f, oparg_f, cam_f, param_def, state_def, shape_wb, vis_wb, link_x, body_reach, cons_i, cons_f, hull_f; } off;

  static void quat_to_mat(const double* q, double* m) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double n = 1.0 / std::sqrt(x * x + y * y + z * z + w * w); x *= n; y *= n; z *= n; w *= n;
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - z * w); m[2] = 2 * (x * z + y * w);
    m[3] = 2 * (x * y + z * w); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - x * w);
    m[6] = 2 * (x * z - y * w); m[7] = 2 * (y * z + x * w); m[8] = 1 - 2 * (x * x + y * y);
  }

  void point(DevScene& d, const int* ib, const float* fb) const {
    d.body_i = ib + off.body_i; d.link_i = ib + off.link_i; d.shape_i = ib + off.shape_i; d.pair_i = ib + off.pair_i;
    d.vis_i = ib + off.vis_i; d.op_i = ib + off.op_i; d.oparg_i = ib + off.oparg_i; d.cam_i = ib + off.cam_i;
    d.dyn_body = ib + off.dyn_body; d.body_plan = ib + off.body_plan; d.frame_slot = ib + off.frame_slot;
    d.link_depth = ib + off.link_depth; d.shape_slot = ib + off.shape_slot;
    d.grp_i = ib + off.grp_i; d.grp_pairs = ib + off.grp_pairs; d.loose_pairs = ib + off.loose_pairs; d.body_reach = fb + off.body_reach;
    d.body_f = fb + off.body_f; d.link_f = fb + off.link_f; d.shape_f = fb + off.shape_f; d.vis_f = fb + off.vis_f;
    d.oparg_f = fb + off.oparg_f; d.cam_f = fb + off.cam_f; d.param_def = fb + off.param_def; d.state_def = fb + off.state_def;
    d.shape_wb = fb + off.shape_wb; d.vis_wb = fb + off.vis_wb; d.link_x = fb + off.link_x;
    d.cons_i = ib + off.cons_i; d.cons_f = fb + off.cons_f; d.hull_f = fb + off.hull_f;
  }

  bool build(const int32_t* ibuf, int ni, const double* fbuf, int nf, int team, int ws_mode = 0, int rs_ashared = 0) {
    if (ni < 2 + 3 * DG_NSECTIONS || ibuf[0] != (int32_t)DG_SCENE_MAGIC || ibuf[1] != DG_NSECTIONS) { error = "bad scene magic / section count"; return false; }
    auto sec_off = [&](int s) { return (size_t)ibuf[2 + 3 * s + 1]; };
    auto sec_len = [&](int s) { return (size_t)ibuf[2 + 3 * s + 2]; };
    for (int s = 0; s < DG_NSECTIONS; s++) {
      bool is_f = ibuf[2 + 3 * s] != 0;
      if (sec_off(s) + sec_len(s) > (size_t)(is_f ? nf : ni)) { error = "section out of range"; return false; }
    }
    const int32_t* hi = ibuf + sec_off(SEC_HDR_I);
    const double* hf = fbuf + sec_off(SEC_HDR_F);
    DevScene& d = dev;
    d = DevScene();
    d.nb = hi[HI_nb]; d.nl = hi[HI_nl]; d.nd = hi[HI_nd]; d.ns = hi[HI_ns]; d.nv = hi[HI_nv]; d.npair = hi[HI_npair];
    d.ncam = hi[HI_ncam]; d.nop = hi[HI_nop]; d.nframes = hi[HI_nframes]; d.S = hi[HI_S]; d.P = hi[HI_P];
    d.substeps = hi[HI_substeps]; d.iters = hi[HI_iterations]; d.maxc = hi[HI_max_contacts]; d.hot_start = hi[HI_hot_start];
    d.ik_iters = hi[HI_ik_iters]; d.n_act = hi[HI_n_act]; d.n_obs = hi[HI_n_obs]; d.n_rew = hi[HI_n_rew]; d.n_term = hi[HI_n_term];
    for (int k = HI_S_BPOS; k <= HI_S_ADDON; k++) d.so[k - HI_S_BPOS] = hi[k];
    for (int k = HI_P_MASS; k <= HI_P_COLOR; k++) d.po[k - HI_P_MASS] = hi[k];
    d.dt = (float)hf[HF_dt]; d.g[0] = (float)hf[HF_gx]; d.g[1] = (float)hf[HF_gy]; d.g[2] = (float)hf[HF_gz];
    d.erp = (float)hf[HF_erp]; d.cerp = (float)hf[HF_contact_erp]; d.slop = (float)hf[HF_linear_slop]; d.margin = (float)hf[HF_contact_margin];
    d.ik_damping = (float)hf[HF_ik_damping]; d.ik_threshold = (float)hf[HF_ik_threshold]; d.max_joint_vel = (float)hf[HF_max_joint_vel];
    d.limit_max_impulse = (float)hf[HF_limit_max_impulse]; d.ik_null_lambda_sq = (float)hf[HF_ik_null_lambda_sq];
    d.team = team;

    ints.clear(); floats.clear();
    auto put_i = [&](int s) { size_t o = ints.size(); ints.insert(ints.end(), ibuf + sec_off(s), ibuf + sec_off(s) + sec_len(s)); ints.push_back(0); return o; };
    auto put_f = [&](int s) { size_t o = floats.size(); const double* p = fbuf + sec_off(s); for (size_t i = 0; i < sec_len(s); i++) floats.push_back((float)p[i]); floats.push_back(0.f); return o; };
    off.body_i = put_i(SEC_BODY_I); off.link_i = put_i(SEC_LINK_I); off.shape_i = put_i(SEC_SHAPE_I); off.pair_i = put_i(SEC_PAIR_I);
    off.vis_i = put_i(SEC_VIS_I); off.op_i = put_i(SEC_OP_I); off.oparg_i = put_i(SEC_OPARG_I); off.cam_i = put_i(SEC_CAM_I);
    off.body_f = put_f(SEC_BODY_F); off.link_f = put_f(SEC_LINK_F); off.shape_f = put_f(SEC_SHAPE_F); off.vis_f = put_f(SEC_VIS_F);
    off.oparg_f = put_f(SEC_OPARG_F); off.cam_f = put_f(SEC_CAM_F); off.param_def = put_f(SEC_PARAM_DEFAULT); off.state_def = put_f(SEC_STATE_DEFAULT);
    off.cons_i = put_i(SEC_CONS_I); off.cons_f = put_f(SEC_CONS_F); off.hull_f = put_f(SEC_HULL_F);
    d.ncons = hi[HI_ncons];

    const int32_t* body_i = ibuf + sec_off(SEC_BODY_I);
    const int32_t* link_i = ibuf + sec_off(SEC_LINK_I);
    const int32_t* shape_i = ibuf + sec_off(SEC_SHAPE_I);
    const double* shape_f = fbuf + sec_off(SEC_SHAPE_F);
    const int32_t* vis_i = ibuf + sec_off(SEC_VIS_I);
    const double* vis_f = fbuf + sec_off(SEC_VIS_F);
    const int32_t* pair_i = ibuf + sec_off(SEC_PAIR_I);
    const int32_t* op_i = ibuf + sec_off(SEC_OP_I);

    // ---- plan tables ----
    std::vector<int> dyn_body, body_plan((size_t)d.nb * BP_W, 0), frame_slot(d.nframes, -1), link_depth(std::max(d.nl, 1), 0), shape_slot(std::max(d.ns, 1), -1);
    int nslot = 0, gv = 0, minv = 0, i0 = 0, nfloat = 0, GD = 1, max_depth = 0, max_nlb = 0, amat = 0;
    for (int b = 0; b < d.nb; b++) {
      const int32_t* bi = body_i + DG_BODY_I_W * b; int* bp = &body_plan[(size_t)b * BP_W];
      int kind = bi[0], l0 = bi[1], nlb = bi[2], ndb = bi[4], baked = bi[7];
      bp[BP_DI] = -1; bp[BP_SLOT] = -1; bp[BP_I0OFF] = -1; bp[BP_UROW] = 2 * bi[3];
      if (kind != 0) {
        bp[BP_DI] = (int)dyn_body.size(); dyn_body.push_back(b);
        int gdim = (kind == 2 ? 6 : 0) + ndb;
        int gs = (gdim + 3) & ~3;   // M^-1 rows are padded to a multiple of 4 floats (zero filled) for 128-bit loads
        bp[BP_GDIM] = gdim; bp[BP_GS] = gs; bp[BP_GVOFF] = gv; bp[BP_MINVOFF] = minv; gv += gs; minv += gdim * gs; GD = std::max(GD, gdim);
        if (kind == 2) { bp[BP_I0OFF] = 36 * nfloat; nfloat++; i0 += 36; }
        // register-resident solver path: at most nrs unit rows (every motor + a couple of active limits)
        int nrs = std::min(2 * ndb, ((ndb + 2 + 3) / 4) * 4); nrs = std::min(((nrs + 3) / 4) * 4, 16);
        bp[BP_NRS] = ndb > 0 ? nrs : 0; bp[BP_AOFF] = amat; amat += bp[BP_NRS] * bp[BP_NRS];
        max_nlb = std::max(max_nlb, nlb);
      }
      if (kind != 0 || !baked) {
        bp[BP_SLOT] = nslot; frame_slot[b] = nslot++;
        for (int k = 0; k < nlb; k++) frame_slot[d.nb + l0 + k] = nslot++;
      }
      int depth_b = 0;
      for (int k = 0; k < nlb; k++) {
        int pl = link_i[DG_LINK_I_W * (l0 + k) + 1];
        link_depth[l0 + k] = pl < 0 ? 0 : link_depth[pl] + 1;
        depth_b = std::max(depth_b, link_depth[l0 + k]);
      }
      bp[BP_DEPTH] = depth_b; max_depth = std::max(max_depth, depth_b);
    }
    int nshw = 0;
    std::vector<float> shape_wb((size_t)std::max(d.ns, 1) * 12, 0.f), vis_wb((size_t)std::max(d.nv, 1) * 12, 0.f);
    for (int s = 0; s < d.ns; s++) {
      if (shape_i[DG_SHAPE_I_W * s + 3]) {
        double m[9]; quat_to_mat(shape_f + DG_SHAPE_F_W * s + 15, m);
        for (int i = 0; i < 9; i++) shape_wb[12 * s + i] = (float)m[i];
        for (int i = 0; i < 3; i++) shape_wb[12 * s + 9 + i] = (float)shape_f[DG_SHAPE_F_W * s + 12 + i];
      } else shape_slot[s] = nshw++;
    }
    for (int s = 0; s < d.nv; s++) if (vis_i[DG_VIS_I_W * s + 2]) {
      double m[9]; quat_to_mat(vis_f + DG_VIS_F_W * s + 19, m);
      for (int i = 0; i < 9; i++) vis_wb[12 * s + i] = (float)m[i];
      for (int i = 0; i < 3; i++) vis_wb[12 * s + 9 + i] = (float)vis_f[DG_VIS_F_W * s + 16 + i];
    }
    std::vector<float> link_x((size_t)std::max(d.nl, 1) * 16, 0.f);
    {
      const double* link_f = fbuf + sec_off(SEC_LINK_F);
      for (int gl = 0; gl < d.nl; gl++) {
        const double* lf = link_f + DG_LINK_F_W * gl; int jt = link_i[DG_LINK_I_W * gl + 2];
        double m[9]; quat_to_mat(lf, m);
        for (int i = 0; i < 9; i++) link_x[16 * gl + i] = (float)m[i];
        const double *a = lf + 10, *dd = lf + 7;
        if (jt == 1) {
          for (int i = 0; i < 3; i++) link_x[16 * gl + 9 + i] = (float)a[i];
          link_x[16 * gl + 12] = (float)(a[1] * dd[2] - a[2] * dd[1]); link_x[16 * gl + 13] = (float)(a[2] * dd[0] - a[0] * dd[2]);
          link_x[16 * gl + 14] = (float)(a[0] * dd[1] - a[1] * dd[0]);
        } else if (jt == 2) for (int i = 0; i < 3; i++) link_x[16 * gl + 12 + i] = (float)a[i];
      }
    }
    // ---- broad-phase groups and body bounding radii ----
    std::vector<float> body_reach(d.nb, 0.f);
    {
      const double* link_f = fbuf + sec_off(SEC_LINK_F);
      std::vector<double> lreach(std::max(d.nl, 1), 0.0);
      auto norm3 = [](const double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
      for (int gl = 0; gl < d.nl; gl++) {
        const double* lf = link_f + DG_LINK_F_W * gl; const int32_t* li = link_i + DG_LINK_I_W * gl;
        double travel = li[2] == 2 ? std::max(std::fabs(lf[20]), std::fabs(lf[21])) : 0.0;
        lreach[gl] = (li[1] < 0 ? 0.0 : lreach[li[1]]) + norm3(lf + 4) + norm3(lf + 7) + travel;
      }
      for (int s = 0; s < d.ns; s++) {
        int b = shape_i[DG_SHAPE_I_W * s], f = shape_i[DG_SHAPE_I_W * s + 1];
        double r = (f < d.nb ? 0.0 : lreach[f - d.nb]) + norm3(shape_f + DG_SHAPE_F_W * s) + shape_f[DG_SHAPE_F_W * s + 11];
        body_reach[b] = std::max(body_reach[b], (float)(r * 1.0001 + 1e-6));
      }
    }
    std::vector<int> grp_i, grp_pairs, loose_pairs;
    {
      std::vector<std::vector<int>> members;   // per group
      std::vector<long long> keys;
      for (int k = 0; k < d.npair; k++) {
        int sa = pair_i[2 * k], sb = pair_i[2 * k + 1];
        bool ba = shape_i[DG_SHAPE_I_W * sa + 3] != 0, bb = shape_i[DG_SHAPE_I_W * sb + 3] != 0;
        int stat = -1, dynb = -1;
        if (ba && !bb) { stat = sa; dynb = shape_i[DG_SHAPE_I_W * sb]; } else if (bb && !ba) { stat = sb; dynb = shape_i[DG_SHAPE_I_W * sa]; }
        if (stat < 0 || body_i[DG_BODY_I_W * dynb] == 0) { loose_pairs.push_back(k); continue; }
        long long key = (long long)stat * d.nb + dynb; size_t g = 0;
        for (; g < keys.size(); g++) if (keys[g] == key) break;
        if (g == keys.size()) { keys.push_back(key); members.emplace_back(); }
        members[g].push_back(k);
      }
      for (size_t g = 0; g < keys.size(); g++) {
        grp_i.push_back((int)(keys[g] / d.nb)); grp_i.push_back((int)(keys[g] % d.nb)); grp_i.push_back((int)grp_pairs.size()); grp_i.push_back((int)members[g].size());
        grp_pairs.insert(grp_pairs.end(), members[g].begin(), members[g].end());
      }
    }
    int GP = 1;
    auto gdim_of_shape = [&](int s) { int b = shape_i[DG_SHAPE_I_W * s]; return body_plan[(size_t)b * BP_W + BP_GDIM]; };
    for (int k = 0; k < d.npair; k++) GP = std::max(GP, gdim_of_shape(pair_i[2 * k]) + gdim_of_shape(pair_i[2 * k + 1]));
    int n_ik = 0;
    for (int k = 0; k < d.nop; k++) n_ik += op_i[DG_OP_I_W * k] == OP_IK_CTRL;
    d.need_react = hi[HI_S_STEP] > hi[HI_S_JREACT];
    d.precise = d.ncons > 0; d.sem = hi[HI_semantics];   // the state row holds reaction wrenches only when a sensor asked for them

    auto put_vi = [&](const std::vector<int>& v) { size_t o = ints.size(); ints.insert(ints.end(), v.begin(), v.end()); ints.push_back(0); return o; };
    auto put_vf = [&](const std::vector<float>& v) { size_t o = floats.size(); floats.insert(floats.end(), v.begin(), v.end()); floats.push_back(0.f); return o; };
    off.dyn_body = put_vi(dyn_body); off.body_plan = put_vi(body_plan); off.frame_slot = put_vi(frame_slot);
    off.link_depth = put_vi(link_depth); off.shape_slot = put_vi(shape_slot);
    off.shape_wb = put_vf(shape_wb); off.vis_wb = put_vf(vis_wb); off.link_x = put_vf(link_x); off.body_reach = put_vf(body_reach);
    off.grp_i = put_vi(grp_i); off.grp_pairs = put_vi(grp_pairs); off.loose_pairs = put_vi(loose_pairs);
    point(d, ints.data(), floats.data());

    d.ndyn = (int)dyn_body.size(); d.nslot = nslot; d.nshw = nshw; d.nfloat = nfloat; d.GD = GD; d.GP = GP;
    if (max_nlb > DG_MINV_MAXL || max_depth + 2 > DG_MINV_MAXD) { error = "a body has more than " + std::to_string((int)DG_MINV_MAXL) + " links or a kinematic tree deeper than " + std::to_string((int)DG_MINV_MAXD - 2); return false; }
    d.max_depth = max_depth; d.max_nlb = max_nlb; d.n_ik = n_ik; d.ngrp = (int)grp_i.size() / 4; d.nloose = (int)loose_pairs.size();

    // ---- workspace layout ----
    // Every region goes either to the HOT workspace (shared memory, offset >= 0) or to the COLD workspace (one
    // global-memory slot per resident team, L1/L2 resident, offset encoded as ~offset < 0).  Inside each memory the
    // articulated-body transients (phase A), the constraint-phase arrays (phase B) and the IK scratch (phase C)
    // alias each other.  ws_mode: 2 = contacts, contact rows and shape poses in shared memory (bodies that rest on
    // contacts), 3 = in the cold workspace (scenes whose bodies rarely touch).
    enum { RC_SMALL = 0, RC_KIN, RC_ABA, RC_SOLVE, RC_CONTACT, RC_SCRATCH };
    auto is_cold = [&](int rc) {
      if (rc == RC_SMALL || rc == RC_SOLVE) return false;                  // header, base state, q / qd, dv, M^-1, unit rows, A: shared
      if (rc == RC_KIN || rc == RC_ABA || rc == RC_SCRATCH) return true;   // frames, joint transforms, ABA transients, scratch: global
      return ws_mode == 3;                                                 // contacts, contact rows, shape poses: per scene
    };
    int fix[2] = {0, 0};                      // persistent part, per memory
    int ph[2][3] = {{0, 0, 0}, {0, 0, 0}};    // phase A / B / C parts, per memory
    // only the contact regions are addressed through the hot/cold selector (negative = cold); the others have a fixed home
    auto enc_rc = [&](int rc, int cold, int off) { return (rc == RC_CONTACT && cold) ? ~off : off; };
    auto take = [&](int rc, int n) { int c = is_cold(rc); int r = fix[c]; fix[c] += (n + 3) & ~3; return std::make_pair(c, r); };
    auto fixed = [&](int rc, int n) { auto pr = take(rc, n); return enc_rc(rc, pr.first, pr.second); };
    d.W_HDR = fixed(RC_SMALL, WH_COUNT);
    d.W_BST = fixed(RC_SMALL, 13 * d.ndyn);
    d.W_DOF = fixed(RC_SMALL, D_COUNT * d.nd);
    d.W_DV = fixed(RC_SMALL, gv);
    d.W_UCNT = fixed(RC_SMALL, d.ndyn);
    d.W_CAPP = fixed(RC_SMALL, 3 * d.maxc);   // accumulated impulse of every contact / friction row (read by the friction bounds)
    d.W_MINV = fixed(RC_SOLVE, minv);
    d.W_KIN = fixed(RC_KIN, 12 * nslot);
    d.W_LINK = fixed(RC_KIN, LK_W * d.nl);
    d.W_I0 = fixed(RC_KIN, i0);
    d.W_EXT = 0; d.W_X = 0;
    // phase regions: offsets are relative to the end of the persistent part of their memory, fixed up below
    struct Pending { int* field; int cold, phase, rel, rc; };
    std::vector<Pending> pend;
    auto phase_take = [&](int* field, int rc, int phase, int n) { int c = is_cold(rc); pend.push_back({field, c, phase, ph[c][phase], rc}); ph[c][phase] += (n + 3) & ~3; };
    d.crow_stride = 2 * (d.GP = GP = (GP + 3) & ~3) + CR_HDR;
    d.ctmp_stride = 1 + 4 * CT_W;
    d.mscr_stride = max_nlb + 6 * (max_depth + 2);
    int gj = 1;
    for (int b = 0; b < d.nb; b++) gj = std::max(gj, (int)body_i[DG_BODY_I_W * b + 4]);
    d.ik_stride = (9 * (max_depth + 2) + 6 * gj + std::max(gj * gj, 36) + 8 * gj + 32 + 3) & ~3;
    phase_take(&d.X_ABA, RC_ABA, 0, AB_W * nslot);
    phase_take(&d.X_LNK, RC_ABA, 0, LX_W * d.nl);
    phase_take(&d.X_I0T, RC_ABA, 0, 36 * nfloat);
    phase_take(&d.X_SHW, RC_CONTACT, 1, 12 * nshw);
    phase_take(&d.X_CON, RC_CONTACT, 1, CT_W * d.maxc);
    phase_take(&d.X_SURV, RC_CONTACT, 1, (d.npair + 31) / 32 + 1);
    phase_take(&d.X_UROW, RC_SOLVE, 1, UR_W * 2 * d.nd);
    phase_take(&d.X_AMAT, RC_SOLVE, 1, amat);
    phase_take(&d.X_CROW, RC_CONTACT, 1, d.crow_stride * 3 * d.maxc);
    phase_take(&d.X_CTMP, RC_SCRATCH, 1, d.ctmp_stride * team);
    phase_take(&d.X_MSCR, RC_SCRATCH, 1, d.mscr_stride * team);
    // row-space team solver: row table + dense A = J M^-1 J^T over all unit and contact rows of the environment
    d.GV = gv;
    // (rows sit at positions of three sections - unit | constraint + normal | friction - each padded to the K rows a lane
    // owns, K <= RS_KMAX: dg_env.cuh "row-space team solver")
    const int rows_max = 2 * d.nd + 3 * d.maxc + 6 * d.ncons;
    auto pad8 = [](int n) { return (n + RS_KMAX - 1) / RS_KMAX * RS_KMAX; };
    const int rows_pad = pad8(2 * d.nd) + pad8(d.maxc + 6 * d.ncons) + pad8(2 * d.maxc);
    d.rs_cap = (team > 1 && gv <= RS_GVMAX) ? std::min(rows_pad, RS_KMAX * team) : 0;
    if (d.ncons > 0 && d.rs_cap < rows_pad) {
      error = "scenes with fixed constraints between models are solved by the row-space team solver: " + std::to_string(rows_max) +
              " rows need a team of at least " + std::to_string((rows_pad + RS_KMAX - 1) / RS_KMAX) + " lanes (and at most " + std::to_string((int)RS_GVMAX) + " generalized coordinates)";
      return false;
    }
    // every environment with contacts is solved in row space (round 1 kept uncoupled ones on the per-body sweeps: its row-space
    // sweeps were slower, profiles/r1_rs_min_sweep.log; DG_RS_MIN restores that for A/B runs)
    // sweep-kernel classes (split schedule): 16 lanes x 2 rows (two environments per warp) and 32 x 2 (one per warp).  The kernel
    // also exists as 8 x 4 (<= 32 positions, four per warp) and 16 x 3 (<= 48): measured on the B200 (profiles/r2_sweep_class_ab.log,
    // DG_SWEEP_CLASSES) 8 x 4 loses everywhere (250 registers, longer steps), 16 x 3 as a middle class wins 1-4 % where few
    // environments need it (r2d2_maze, ur_robotiq) and loses 4 % where all do (from_the_readme) - so it is off by default.
    d.rs_cls_k[0] = 2; d.rs_cls_r[0] = 32; d.rs_cls_k[1] = 0; d.rs_cls_r[1] = 0; d.rs_cls_k[2] = 2; d.rs_cls_r[2] = 64;
    d.solver = 1; d.rs_min = 0;   // every environment with contacts is solved in row space (round 1 kept uncoupled ones on the per-body sweeps: its row-space sweeps were slower, profiles/r1_rs_min_sweep.log; DG_RS_MIN restores that for A/B runs)
    d.rs_ashared = 0; d.X_RSAS = 0; (void)rs_ashared;   // (shared-memory home of A: measured slower, removed)
    phase_take(&d.X_RSA, RC_SCRATCH, 1, d.rs_cap * d.rs_cap + RS_KMAX * team);
    phase_take(&d.X_RSV, RC_SCRATCH, 1, d.rs_cap * 2 * gv);      // per row: J and M^-1 J^T, dense over all bodies' coordinates
    phase_take(&d.X_RSREC, RC_SCRATCH, 1, d.rs_cap * RR_W);
    phase_take(&d.X_IK, RC_SCRATCH, 2, n_ik ? d.ik_stride * std::min(team, n_ik) : 0);
    for (auto& pd : pend) *pd.field = enc_rc(pd.rc, pd.cold, fix[pd.cold] + pd.rel);
    d.w_total = fix[0] + std::max(ph[0][0], std::max(ph[0][1], ph[0][2]));
    d.g_total = fix[1] + std::max(ph[1][0], std::max(ph[1][1], ph[1][2]));
    d.ws_mode = ws_mode;
    return true;
  }
};

}  // namespace dg
#endif
