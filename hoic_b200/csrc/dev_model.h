// dev_model.h -- flat model constants as the kernels see them (float32, fixed maxima) + per-env smem layout.
// Built on the host from the blob (hoic_capi.cu: build_dev_model), lives in global memory, read through the
// read-only path.  Sizes are those of the HandMimic model class: nv = 32 = one lane per dof.
#pragma once
#include <stdint.h>

#define HN 32          // lanes = max bodies / dofs / geoms / joints
#define H_NQ 36
#define H_MAXCAND 128
#ifndef H_NCON
#define H_NCON 24      // contact capacity per env (<= HOIC_MAX_CONTACTS)
#endif
#ifndef H_NBASIS
#define H_NBASIS 56    // contact Jacobian basis rows per env (condim 3: 3 per contact, condim 4: 4)
#endif
#ifndef H_NEDGE
#define H_NEDGE 96     // pyramid edge rows per env
#endif
#define H_RFR 0        // row index ranges: friction-loss rows [0,32)
#define H_RLIM 32      //                   joint-limit rows   [32,64)
#define H_RCON 64      //                   contact edge rows  [64,64+H_NEDGE)
#define H_NROW (H_RCON + H_NEDGE)
#define H_OBS 617
#define H_FRAME 212
#define H_NDOP 13      // directions of the mesh 26-DOP (pair culling)
#define H_NHG 19       // hand collision geoms tracked for contact averaging (ho_im4.py:80-87)

enum { G_PLANE = 0, G_SPHERE = 2, G_CAPSULE = 3, G_BOX = 6, G_MESH = 7 };
enum { J_FREE = 0, J_SLIDE = 2, J_HINGE = 3 };
enum { K_PLANE_CAPSULE = 0, K_PLANE_BOX, K_PLANE_MESH, K_CAP_CAP, K_CAP_BOX, K_BOX_BOX, K_CONVEX };

struct DevCand {
  int g1, g2, kind, condim;     // g1/g2 already ordered by geom type (then id), as mj_collideGeoms does
  float margin, includemargin;
  float friction[3];            // f0 (both tangents), f1 (torsion), f2 (roll)
  float solref[2], solimp[5];
  float dA_tran, dA_rot;        // body_invweight0 sums of the two bodies
  int hand_slot;                // >=0: g1 is hand geom hand_geom0+slot and g2 an object collision geom (record_contact)
  unsigned chain1, chain2;      // ancestor-dof masks of the two geoms' bodies
  int pad;
};

struct DevModel {
  int nq, nv, nbody, ngeom, njnt, nu, ncand, nlevel;
  int iterations, mpr_iterations, nhand_q, nhand_v, nroot, sim_step, w_size, n_hand_body;
  int obj_body, palm_body, hand_geom0, hand_geom1, obj_geom0, obj_geom1, last_geom, residual_force;
  int surface_contact, explain_force, org_qadr, pad0;
  float timestep, tolerance, impratio, mpr_tolerance, meaninertia, gravity[3];
  float hand_mass, rf_scale, rt_scale, pos_thresh, rot_thresh, jpos_thresh, obj_pos_thresh, obj_rot_thresh;
  float obj_mass, obj_inertia[3];
  // bodies
  int body_parent[HN], body_level[HN], body_rootslot[HN], body_dofadr[HN], body_dofnum[HN], body_jntadr[HN],
      body_jntnum[HN], body_firstchild[HN], body_nextsib[HN], body_weld[HN];
  unsigned body_chain[HN];       // bit i set: dof i moves this body
  float body_pos[HN][3], body_quat[HN][4], body_ipos[HN][3], body_imat[HN][9], body_mass[HN], body_inertia[HN][3];
  int root_body[4];
  float root_invmass[4];
  // joints
  int jnt_type[HN], jnt_body[HN], jnt_qposadr[HN], jnt_dofadr[HN], jnt_limited[HN];
  float jnt_pos[HN][3], jnt_axis[HN][3], jnt_range[HN][2], jnt_margin[HN], jnt_solref[HN][2], jnt_solimp[HN][5];
  float jnt_invweight[HN], qpos0[H_NQ];
  // dofs
  int dof_body[HN], dof_jnt[HN], dof_parent[HN], dof_rootslot[HN], dof_dotzero[HN];
  unsigned dof_anc[HN];          // strict ancestor dofs
  unsigned dof_premask[HN];      // dofs whose velocity is accumulated before cdof_dot of this dof is formed (mj_comVel)
  float dof_armature[HN], dof_damping[HN], dof_frictionloss[HN], fr_D[HN], fr_B[HN], fr_Rf[HN];
  // dof-tree tables of the tree-sparse L'DL solver (dev_math.cuh: tree_solve).  Supported topology: chains plus at most one
  // "hub" dof with several child chains (the palm's last dof); checked in build_dev_model.
  int dof_depth[HN];             // number of ancestor dofs
  unsigned dof_anc_at[HN][3];    // byte d: ancestor dof at depth d (d < depth), 0xff otherwise
  unsigned dof_desc_at[HN][3];   // byte l: the unique descendant dof at depth l inside this dof's chain, 0xff if none / hub fan-out
  int dof_is_trunk[HN];          // hub dof or one of its ancestors
  unsigned dof_desc[HN];         // all descendant dofs
  unsigned branch_mask;          // dofs below the hub
  unsigned dof_depth_nib[4];     // 4-bit depth of dof j at bit 4*(j%8) of word j/8
  int hub_depth, tree_maxdepth, branch_lo, branch_n;   // branch dofs are the contiguous range [branch_lo, branch_lo + branch_n)
  // geoms
  int geom_type[HN], geom_body[HN], geom_vertadr[HN], geom_vertnum[HN];
  float geom_pos[HN][3], geom_mat[HN][9], geom_size[HN][3], geom_rbound[HN];
  const float4* mesh_vert;       // device pointer: xyz + original vertex index (int bits); every mesh is stored as clusters of 32
                                 // spatially close vertices (kd split at model build), cluster k of geom g = vertices [32k, 32k+32)
  const float4* mesh_cl;         // device pointer: per cluster the local AABB (lo, hi) -> support-function pruning
  int geom_cladr[HN], geom_ncl[HN];
  float geom_aabb_c[HN][3], geom_aabb_h[HN][3];   // mesh geoms: local AABB centre / half sizes (conservative pair culling)
  float geom_dop[HN][H_NDOP][2];                  // mesh geoms: [min, max] of the vertices along the c_dop directions
  // env
  int hand_body_idx[HN];
  float base_pose[HN], ctrl_scale[HN], jkp[HN], jkd[HN], torque_lim[HN];
  int stage_bytes, pad1;         // k_step stages this many leading bytes of the model into shared memory (0: reads stay global)
  // collision candidates in MuJoCo's signature order; LAST member: only the first ncand entries are staged
  DevCand cand[H_MAXCAND];
};

// ---------------------------------------------------------------------------------------------------------------
// per-environment shared-memory layout (float offsets).  Persistent part survives a substep boundary (the env logic
// reads the PREVIOUS forward's M / bias / cdof / poses: SURVEY.md S1); the union part is reused inside a substep.
#define S_QPOS 0
#define S_QVEL (S_QPOS + H_NQ)
#define S_WARM (S_QVEL + HN)
#define S_QACC (S_WARM + HN)
#define S_CTRL (S_QACC + HN)
#define S_QAPP (S_CTRL + HN)
#define S_ACT (S_QAPP + HN)
#define S_ORG (S_ACT + HN)             // 4
#define S_XPOS (S_ORG + 4)             // 32*3, relative to S_ORG
#define S_XQUAT (S_XPOS + 96)          // 32*4
#define S_GPOS (S_XQUAT + 128)         // 32*3
#define S_GMAT (S_GPOS + 96)           // 32*9
#define S_COM (S_GMAT + 288)           // 4 roots * 3 -> 12
#define S_CDOF (S_COM + 12)            // 32*6
#define S_BIAS (S_CDOF + 192)
#define S_SMOOTH (S_BIAS + HN)         // qfrc_smooth
#define S_ASMOOTH (S_SMOOTH + HN)      // qacc_smooth
#define S_MA (S_ASMOOTH + HN)
#define S_MV (S_MA + HN)
#define S_GRAD (S_MV + HN)
#define S_SEARCH (S_GRAD + HN)
#define S_TMP (S_SEARCH + HN)          // 32
#define S_FCON (S_TMP + HN)            // qfrc_constraint
#define S_MISC (S_FCON + HN)           // 16 scalars: ncon, nlim, nbasis, nedge, flags, ...
#define MS_STRIDE 12                   // row stride of S_MS (16-byte aligned rows; entries beyond the depth stay zero)
#define S_MS (S_MISC + 16)             // 32*12: M by (dof, ancestor depth): entry d of row i = M[i, ancestor of i at depth d]
#define S_H (S_MS + 384)               // 576: lower triangle of the Newton Hessian / its Cholesky factor in 4x4 blocks
                                       //      (dev_math.cuh: warp_chol_solve); scratch of tree_solve outside the Newton iterations
// S_MS and S_H sit right below the union: both are dead after the last substep, so the residual-force QP of the step
// tail starts its scratch at S_MS (QP_BASE) and the union only has to hold the position and constraint stages.
// The per-hand-geom contact sums of record_contact (19 x 12 + counts) live in global memory (BatchDev.csum): they are
// touched once per recorded contact, and every float of shared memory per env counts towards the warps an SM can hold.
#define S_UNION (S_H + 576)
#define QP_BASE S_MS
#define CSUM_STRIDE 272                // global: 19*12 contact sums (240) + 32 counts
// position / velocity stage temporaries
#define U_XMAT (S_UNION)               // 32*9
#define U_XAXIS (U_XMAT + 288)         // 32*3 per joint
#define U_XANCH (U_XAXIS + 96)
#define U_CINERT (U_XANCH + 96)        // 32*10
#define U_CRB (U_CINERT + 320)
#define U_CDOFDOT (U_CRB + 320)        // 32*6
#define U_CVEL (U_CDOFDOT + 192)
#define U_CACC (U_CVEL + 192)
#define U_CFRC (U_CACC + 192)
#define U_POS_END (U_CFRC + 192)
// constraint stage (aliases the above; built after the velocity stage)
#define U_CON (S_UNION)                // H_NCON * 16: pos3 frame9 dist cand(int) 2 pad
#define B_STRIDE 17                    // a contact Jacobian row touches <= 16 dofs (one hand chain + the object, or two
                                       // finger chains): columns are stored compacted in dof order, see build_rows
#define U_B (U_CON + H_NCON * 16)      // H_NBASIS * B_STRIDE basis rows
#define U_AREF (U_B + H_NBASIS * B_STRIDE)   // H_NROW each:
#define U_D (U_AREF + H_NROW)
#define U_JAR (U_D + H_NROW)
#define U_JV (U_JAR + H_NROW)
#define U_FORCE (U_JV + H_NROW)
#define U_RINFO (U_FORCE + H_NROW)     // int per row: limit rows dof | sign<<8 ; edge rows contact | edge<<8
#define U_BV (U_RINFO + H_NROW)        // H_NBASIS: basis . v products
#define U_BMASK (U_BV + H_NBASIS)      // H_NBASIS: dof mask of each basis row (bit i: dof i has a column)
#define U_CINFO (U_BMASK + H_NBASIS)      // H_NCON * 8: basis offset, nbasis, edge offset, nedge, mu0, mu1, D, pad
#define U_MPRQ (U_CINFO + H_NCON * 8)  // H_MAXCAND 16-bit indices: queued MPR candidates (later H_NBASIS ints: U_BINFO)
#define U_CON_END (U_MPRQ + (H_MAXCAND / 2 > H_NBASIS ? H_MAXCAND / 2 : H_NBASIS))
#define U_QP_END (QP_BASE + 3168)      // residual-force QP scratch (dev_env.cuh: QP_*; static_assert there)
#define U_TS_END (S_UNION + 32 * 21)   // tree_solve's branch products (dev_math.cuh: TS_P)
#define S_MAX2(a, b) ((a) > (b) ? (a) : (b))
#define S_TOTAL ((S_MAX2(S_MAX2(U_CON_END, U_POS_END), S_MAX2(U_QP_END, U_TS_END)) + 3) & ~3)

// misc slots
#define MI_NCON 0
#define MI_NLIM 1
#define MI_NBASIS 2
#define MI_NEDGE 3
#define MI_FLAGS 4      // bit0 physics failure (NaN / huge), bit1 capacity overflow
#define MI_ITER 5

// per-env persistent state in global memory (floats), AoS so that a warp reads it coalesced
#define ST_QPOS 0
#define ST_QVEL 36
#define ST_WARM 68
#define ST_PQPOS 100     // state of the previous forward (stale derived quantities are recomputed from it)
#define ST_PQVEL 136
#define ST_INTS 168      // cur_t, motion_id, start_idx, expert_len, flags
#define ST_RFC 176       // rfc_score, rest_force3, rest_torque3, n contact geoms
#define ST_STRIDE 192
