// dev_model.h -- flat model constants as the kernels see them (float32, fixed maxima) + per-env smem layout.
// Built on the host from the blob (hoic_capi.cu: build_dev_model), lives in global memory, read through the
// read-only path.  Sizes are those of the HandMimic model class: nv = 32 = one lane per dof.
#pragma once
#include <stdint.h>

#define HN 32          // lanes = max bodies / dofs / geoms / joints
#define H_NQ 36
#define H_MAXCAND 128
#ifndef H_NCON
#define H_NCON 32      // contact capacity per env (<= HOIC_MAX_CONTACTS)
#endif
#ifndef H_NBASIS
#define H_NBASIS 80    // contact Jacobian basis rows per env (condim 3: 3 per contact, condim 4: 4)
#endif
#ifndef H_NEDGE
#define H_NEDGE 128    // pyramid edge rows per env
#endif
// capacities of the GLOBAL working set (BatchDev.scratch): what k_p2 produces.  A kernel compiled with smaller H_NCON /
// H_NBASIS / H_NEDGE (k_p3's small tier) holds a smaller constraint stage in shared memory and only takes envs that fit.
#define HG_NCON 32
#define HG_NBASIS 80
#define HG_NEDGE 128
// capacity of the solver's SMALL tier (k_p3s / k_as): envs whose constraint stage fits are solved there, the rest by k_p3 / k_a
#ifndef P3S_NCON
#define P3S_NCON 8
#define P3S_NBASIS 24
#define P3S_NEDGE 32
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
// Per-environment working-set layout (float offsets).  One table describes every field once; each kernel of the step
// pipeline (k_p1 / k_p2 / k_p3 / k_tail, see hoic_step.cuh) is compiled in its own translation unit with H_KERNEL set to
// its bit, and gets a COMPACT shared-memory layout holding only the fields it touches (offsets S_*: a constexpr prefix
// sum over the fields whose `use` mask contains the kernel).  Between kernels the working set lives in global memory
// (BatchDev.scratch, L2-resident) in the FULL layout (offsets G_*).  The single-warp kernels of hoic_capi.cu (reset,
// debug entries, preprocessing, standalone QP) use the full layout in shared memory as well.
// "Persistent" fields survive a substep boundary (the env logic reads the PREVIOUS forward's M / bias / cdof / poses:
// SURVEY.md S1); the union region at the end is reused inside a substep.
#define HK_P1 1u        // Euler of the previous substep, stable PD, J^T wrenches, position + velocity stage, qacc_smooth
#define HK_P2 2u        // collision, contact recording, constraint rows
#define HK_P3 4u        // Newton solve
#define HK_TAIL 8u      // last Euler, averages, residual-force QP, termination, reward, observation
#define HK_FULL 0xffu
#ifndef H_KERNEL
#define H_KERNEL HK_FULL
#endif
#define MS_STRIDE 12                   // row stride of S_MS (16-byte aligned rows; entries beyond the depth stay zero)
#define B_STRIDE 17                    // a contact Jacobian row touches <= 16 dofs (one hand chain + the object, or two
                                       // finger chains): columns are stored compacted in dof order, see build_rows
// sizes of the three users of the union region (floats)
#define UN_POS (288 + 96 + 96 + 320 + 320)                                             /* position / velocity stage */
#define UN_CON_OF(nc, nb, ne) ((nc) * 16 + (nb) * B_STRIDE + 6 * (H_RCON + (ne)) + 2 * (nb) + (nc) * 8 + (H_MAXCAND / 2 > (nb) ? H_MAXCAND / 2 : (nb)))
// k_p2 builds the rows but never touches the solver's per-row work arrays (JAR, JV, FORCE, BV), which lie at the end
#define UN_CON_BUILD_OF(nc, nb, ne) (UN_CON_OF(nc, nb, ne) - 3 * (H_RCON + (ne)) - (nb))
#define HK_P12 (HK_P1 | HK_P2)  // fused kernel k_p12 (k_p12.cu): k_p1's and k_p2's work of a substep in one launch
#define UN_CON ((H_KERNEL) == HK_P2 || (H_KERNEL) == HK_P12 ? UN_CON_BUILD_OF(H_NCON, H_NBASIS, H_NEDGE) : UN_CON_OF(H_NCON, H_NBASIS, H_NEDGE))
#define UN_CON_G UN_CON_OF(HG_NCON, HG_NBASIS, HG_NEDGE)
#define UN_QP (3168 - 384 - 576)       /* residual-force QP scratch starts at S_MS (QP_BASE) and runs over S_H into the union */
#define UN_TS (32 * 21)                /* tree_solve's branch products (dev_math.cuh: TS_P) */
#define UN_TSF 480                     /* tree_solve's factor scratch (TS_LS .. TS_T) when it lives in the union (k_p1) */
#define S_MAX2(a, b) ((a) > (b) ? (a) : (b))
#define HK_A (HK_P1 | HK_P3)   // fused kernel k_a (k_a.cu): Newton solve of substep k, then k_p1's work for substep k + 1
#define UN_SIZE ((H_KERNEL) == HK_P1 ? S_MAX2(UN_POS, UN_TS + UN_TSF) : ((H_KERNEL) == HK_P2 || (H_KERNEL) == HK_P3 || (H_KERNEL) == HK_P12) ? UN_CON : \
                 (H_KERNEL) == HK_A ? S_MAX2(S_MAX2(UN_POS, UN_TS), UN_CON) : \
                 (H_KERNEL) == HK_TAIL ? S_MAX2(UN_QP, UN_TS) : S_MAX2(S_MAX2(UN_POS, UN_CON), S_MAX2(UN_QP, UN_TS)))
#define UN_FULL S_MAX2(S_MAX2(UN_POS, UN_CON_G), S_MAX2(UN_QP, UN_TS))

#define H_FIELDS(X)                                                                                                      \
  X(QPOS, H_NQ, HK_P1 | HK_P2 | HK_TAIL)                                                                                   \
  X(QVEL, HN, HK_P1 | HK_P2 | HK_TAIL)                                                                                     \
  X(WARM, HN, HK_P1 | HK_P3 | HK_TAIL)                                                                                     \
  X(QACC, HN, HK_P1 | HK_P3 | HK_TAIL)                                                                                     \
  X(CTRL, HN, HK_P1)                                                                                                       \
  X(QAPP, HN, HK_P1)                                                                                                       \
  X(ACT, HN, HK_P1)                                                                                                        \
  X(ORG, 4, HK_P1 | HK_P2 | HK_TAIL)                  /* hand-root translation the stored positions are relative to */     \
  X(XPOS, 96, HK_P1 | HK_TAIL)                                                                                             \
  X(XQUAT, 128, HK_P1 | HK_TAIL)                                                                                           \
  X(GPOS, 96, HK_P2 | HK_TAIL)                                                                                             \
  X(GMAT, 288, HK_P2 | HK_TAIL)                                                                                            \
  X(COM, 12, HK_P1 | HK_P2)                           /* 4 roots * 3 */                                                    \
  X(CDOF, 192, HK_P1 | HK_P2)                                                                                              \
  X(BIAS, HN, HK_P1)                                                                                                       \
  X(SMOOTH, HN, HK_P1 | HK_P3 | HK_TAIL)              /* qfrc_smooth */                                                    \
  X(ASMOOTH, HN, HK_P1 | HK_P3)                       /* qacc_smooth */                                                    \
  X(MA, HN, HK_P3)                                                                                                         \
  X(MV, HN, HK_P3)                                                                                                         \
  X(GRAD, HN, HK_P3)                                                                                                       \
  X(SEARCH, HN, HK_P3)                                                                                                     \
  X(FCON, HN, HK_P1 | HK_P3 | HK_TAIL)                /* qfrc_constraint */                                                \
  X(MISC, 32, HK_FULL)                                /* int scalars: counts, flags, per-step counters (MI_*) */           \
  X(MS, 384, HK_P1 | HK_P3 | HK_TAIL)                 /* M by (dof, ancestor depth): entry d of row i = M[i, anc_d(i)] */   \
  X(H, 576, HK_P3 | HK_TAIL)                          /* Newton Hessian / Cholesky factor in 4x4 blocks; tree_solve scratch */ \
  X(UNION, UN_SIZE, HK_FULL)

enum {
#define X(name, size, use) FI_##name,
  H_FIELDS(X)
#undef X
  FI_COUNT
};
struct HField { int size; unsigned use; };
#define H_POISON (1 << 24)     // offset of a field the kernel's layout does not hold: any access faults
#define P12_P1_FIELDS 1152     // k_p12: union floats of k_p1's own stages (max(UN_POS, UN_TS + UN_TSF)); its private fields follow
__host__ __device__ constexpr int h_field_off(int f, unsigned kernel, int union_size) {
  const HField tab[FI_COUNT] = {
#define X(name, size, use) {FI_##name == FI_UNION ? union_size : (size), (use)},
      H_FIELDS(X)
#undef X
  };
  if (kernel == HK_P12) {
    // k_p12: the fields k_p2 reads are laid out first, then the union at k_p2's size; the fields only k_p1's half touches
    // (dead once that half has stored its results) live INSIDE the union, behind the 1152 floats k_p1's own stages use
    int common = 0, p1only = 0, off_common = 0, off_p1 = 0;
    for (int i = 0; i < FI_COUNT; i++) {
      if (i == FI_UNION) continue;
      const bool c = (tab[i].use & HK_P2) != 0, q = !c && (tab[i].use & HK_P1) != 0;
      if (i == f) { off_common = common; off_p1 = p1only; }
      if (c) common += tab[i].size;
      if (q) p1only += tab[i].size;
    }
    if (f >= FI_COUNT) return common + union_size;
    if (f == FI_UNION) return common;
    if (tab[f].use & HK_P2) return off_common;
    if (tab[f].use & HK_P1) return common + P12_P1_FIELDS + off_p1;
    return H_POISON;
  }
  int o = 0;
  for (int i = 0; i < f && i < FI_COUNT; i++)
    if (tab[i].use & kernel) o += tab[i].size;
  if (f < FI_COUNT && !(tab[f].use & kernel)) return H_POISON;
  return o;
}
__host__ __device__ constexpr int h_field_size(int f) {
  const HField tab[FI_COUNT] = {
#define X(name, size, use) {(size), (use)},
      H_FIELDS(X)
#undef X
  };
  return tab[f].size;
}
#define S_OFF(name) h_field_off(FI_##name, H_KERNEL, UN_SIZE)
#define G_OFF(name) h_field_off(FI_##name, HK_FULL, UN_FULL)
#define F_SIZE(name) h_field_size(FI_##name)
#define S_TOTAL ((h_field_off(FI_COUNT, H_KERNEL, UN_SIZE) + 3) & ~3)
#define G_TOTAL ((h_field_off(FI_COUNT, HK_FULL, UN_FULL) + 3) & ~3)

#define S_QPOS S_OFF(QPOS)
#define S_QVEL S_OFF(QVEL)
#define S_WARM S_OFF(WARM)
#define S_QACC S_OFF(QACC)
#define S_CTRL S_OFF(CTRL)
#define S_QAPP S_OFF(QAPP)
#define S_ACT S_OFF(ACT)
#define S_ORG S_OFF(ORG)
#define S_XPOS S_OFF(XPOS)             // 32*3, relative to S_ORG
#define S_XQUAT S_OFF(XQUAT)           // 32*4
#define S_GPOS S_OFF(GPOS)             // 32*3
#define S_GMAT S_OFF(GMAT)             // 32*9
#define S_COM S_OFF(COM)
#define S_CDOF S_OFF(CDOF)             // 32*6
#define S_BIAS S_OFF(BIAS)
#define S_SMOOTH S_OFF(SMOOTH)
#define S_ASMOOTH S_OFF(ASMOOTH)
#define S_MA S_OFF(MA)
#define S_MV S_OFF(MV)
#define S_GRAD S_OFF(GRAD)
#define S_SEARCH S_OFF(SEARCH)
#define S_FCON S_OFF(FCON)
#define S_MISC S_OFF(MISC)
#define S_MS S_OFF(MS)
#define S_H S_OFF(H)
// S_MS and S_H sit right below the union: both are dead after the last substep, so the residual-force QP of the step
// tail starts its scratch at S_MS (QP_BASE) and the union only has to hold what the kernel's own stages need.
// The per-hand-geom contact sums of record_contact (19 x 12 + counts) live in global memory (BatchDev.csum).
#define S_UNION S_OFF(UNION)
#define QP_BASE S_MS
#define CSUM_STRIDE 272                // global: 19*12 contact sums (240) + 32 counts
// position / velocity stage temporaries
#define U_XMAT (S_UNION)               // 32*9
#define U_XAXIS (U_XMAT + 288)         // 32*3 per joint
#define U_XANCH (U_XAXIS + 96)
#define U_CINERT (U_XANCH + 96)        // 32*10
#define U_CRB (U_CINERT + 320)
#define U_POS_END (U_CRB + 320)
// velocity-stage temporaries lie over position-stage ones that are dead by then (U_XMAT / U_XAXIS / U_XANCH, U_CRB)
#define U_DV (U_XMAT)                  // 32*6: cdof * qvel
#define U_DA (U_XMAT + 192)            // 32*6: cdof_dot * qvel
#define U_CFRC (U_CRB)                 // 32*6
#define U_JQ (U_CRB)                   // 32*4: hinge rotation quaternion of each joint (first half of pos_stage only)
// constraint stage (aliases the above; built after the velocity stage); offsets relative to the union start
#define UL_CON(nc, nb, ne) 0                                     // nc * 16: pos3 frame9 dist cand(int) 2 pad
#define UL_B(nc, nb, ne) ((nc) * 16)                             // nb * B_STRIDE basis rows
#define UL_AREF(nc, nb, ne) (UL_B(nc, nb, ne) + (nb) * B_STRIDE) // (H_RCON + ne) each:
#define UL_D(nc, nb, ne) (UL_AREF(nc, nb, ne) + H_RCON + (ne))
#define UL_RINFO(nc, nb, ne) (UL_D(nc, nb, ne) + H_RCON + (ne))  // int per row: limit rows dof | sign<<8 ; edge rows contact | edge<<8
#define UL_BMASK(nc, nb, ne) (UL_RINFO(nc, nb, ne) + H_RCON + (ne))   // nb: dof mask of each basis row (bit i: dof i has a column)
#define UL_CINFO(nc, nb, ne) (UL_BMASK(nc, nb, ne) + (nb))       // nc * 8: basis offset, nbasis, edge offset, nedge, mu0, mu1, D, pad
#define UL_MPRQ(nc, nb, ne) (UL_CINFO(nc, nb, ne) + (nc) * 8)    // H_MAXCAND 16-bit indices: queued MPR candidates (later nb ints: U_BINFO)
// ... the solver's work arrays (absent from k_p2's layout)
#define UL_JAR(nc, nb, ne) (UL_MPRQ(nc, nb, ne) + (H_MAXCAND / 2 > (nb) ? H_MAXCAND / 2 : (nb)))
#define UL_JV(nc, nb, ne) (UL_JAR(nc, nb, ne) + H_RCON + (ne))
#define UL_FORCE(nc, nb, ne) (UL_JV(nc, nb, ne) + H_RCON + (ne))
#define UL_BV(nc, nb, ne) (UL_FORCE(nc, nb, ne) + H_RCON + (ne)) // nb: basis . v products
#define U_S(f) (S_UNION + UL_##f(H_NCON, H_NBASIS, H_NEDGE))     // in this kernel's shared memory
#define U_G(f) (G_OFF(UNION) + UL_##f(HG_NCON, HG_NBASIS, HG_NEDGE))   // in the global working set
#define U_CON U_S(CON)
#define U_B U_S(B)
#define U_AREF U_S(AREF)
#define U_D U_S(D)
#define U_JAR U_S(JAR)
#define U_JV U_S(JV)
#define U_FORCE U_S(FORCE)
#define U_RINFO U_S(RINFO)
#define U_BV U_S(BV)
#define U_BMASK U_S(BMASK)
#define U_CINFO U_S(CINFO)
#define U_MPRQ U_S(MPRQ)
#define U_CON_END (S_UNION + UN_CON)
static_assert(UL_BV(H_NCON, H_NBASIS, H_NEDGE) + H_NBASIS == UN_CON_OF(H_NCON, H_NBASIS, H_NEDGE), "constraint stage layout");
#define U_QP_END (QP_BASE + 3168)      // residual-force QP scratch (dev_env.cuh: QP_*; static_assert there)
#define U_TS_END (S_UNION + UN_TS)
static_assert(UN_POS == 1120 && (UN_CON % 4) == 0 && (H_NBASIS * B_STRIDE) % 4 == 0 && H_NROW % 4 == 0 && H_NCON % 4 == 0 &&
              H_NBASIS % 4 == 0 && H_NCON <= HG_NCON && H_NBASIS <= HG_NBASIS && H_NEDGE <= HG_NEDGE, "union sub-layouts");

// misc slots
#define MI_NCON 0
#define MI_NLIM 1
#define MI_NBASIS 2
#define MI_NEDGE 3
#define MI_FLAGS 4      // bit0 physics failure (NaN / huge), bit1 capacity overflow
#define MI_ITER 5
// per-env counters of the running step (summed into hoic_counters by the step tail)
#define MI_C_SUBSTEPS 8
#define MI_C_NCON 9
#define MI_C_NEFC 10
#define MI_C_NEWTON 11
#define MI_C_LS 12
#define MI_C_MPRCALLS 13
#define MI_C_MPRITERS 14
#define MI_C_MAXNCON 15
#define MI_C_MAXNBASIS 16
#define MI_C_MAXNEDGE 17
#define MI_C_FAILPHYS 18
#define MI_C_FAILOVER 19
#define MI_CARRIED 20    // this step started from the working set the previous step left (no pass -1 recomputation)
// a Newton solve that a capped launch left open (k_p3s with an iteration budget, continued by k_p3c; hoic_step "carry" mode)
#define MI_SV_OPEN 21    // 1: the solve of this substep is open, QACC / MA / JAR in the working set hold its state
#define MI_SV_ITER 22    // iterations done so far
#define MI_SV_COST 23    // float bits: cost, cost before the last move
#define MI_SV_OLDCOST 24

// per-env persistent state in global memory (floats), AoS so that a warp reads it coalesced
#define ST_QPOS 0
#define ST_QVEL 36
#define ST_WARM 68
#define ST_PQPOS 100     // state of the previous forward (stale derived quantities are recomputed from it)
#define ST_PQVEL 136
#define ST_INTS 168      // cur_t, motion_id, start_idx, expert_len, flags, work, episode counter, working set valid
#define ST_RFC 176       // rfc_score, rest_force3, rest_torque3, n contact geoms
#define ST_STRIDE 192
