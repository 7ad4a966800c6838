/* rdc.h -- C ABI of the B200-native rdcFEs hot path (FE assembly of the RDC operators + Krylov solve).
 *
 * This is the drop-in boundary.  Every entry point names the
 * reference interface it replaces (file:line into InSilicoModellingGroup/rdcFEs):
 *
 *   rdc_create / rdc_set_*      <- the one-time hand-over the assemble_* callbacks would do after es.init()
 *                                  (adpm.C:43, pihna.C es.init(), ripf.C:47, ...): flattened mesh + dof map
 *   rdc_set_params              <- es.parameters reads at adpm.C:364-414, pihna.C:358-381, ripf.C:377-408,
 *                                  proteas.C:376-410, coupled_hcc.C:450-461
 *   rdc_set_elem_field          <- "Tracts" CONSTANT MONOMIAL system (adpm.C:32-37, 453-458)
 *   rdc_set_nodal_field         <- "RT" system (ripf.C:36-41, 473-478), "AUX" system (proteas.C:470-482)
 *   rdc_update_coords           <- SolidSystem::update() -> mesh_position_set (solid_system.C:103-108)
 *   rdc_rotate                  <- *older = *old; *old = *current  (adpm.C:71-72 and the 4 siblings)
 *   rdc_assemble                <- matrix->zero(); rhs->zero(); assemble_<model>(es, name)
 *                                  (adpm.C:324-652, pihna.C:318-758, ripf.C:337-673, proteas.C:338-705,
 *                                   coupled_hcc.C:414-649)
 *   rdc_solve                   <- LinearSolver<Number>::solve(K,u,F,tol,maxits) inside
 *                                  TransientLinearImplicitSystem::solve() (adpm.C:74, pihna.C:80, ripf.C:83,
 *                                  proteas.C:78, coupled_hcc.C:114)
 *   rdc_clamp                   <- check_solution (adpm.C:654-688, pihna.C:760-803, ripf.C:675-775,
 *                                  proteas.C:707-750, coupled_hcc.C:695-731)
 *   rdc_step                    <- one iteration of the time loop body (adpm.C:63-76)
 *   rdc_get_solution            <- es.build_solution_vector (paraview.h:24-25) at output steps
 *   rdc_get_solution_owned      <- the rank-local part of system.solution (distributed PETSc vector)
 *   rdc_set_subdomains /
 *   rdc_region_volumes /
 *   rdc_region_last_mean        <- the element loops of save_solution (adpm.C:690-829, pihna.C:842-976,
 *                                  ripf.C:777-864) behind the per-case CSV output
 *   rdc_download_csr            <- parity only: the assembled PETSc AIJ matrix and rhs
 *
 * Rules: plain C types only; every function returns 0 (RDC_OK) or a negative RDC_E* code; the message
 * for the last failure on a context is available from rdc_last_error(); host input buffers are copied
 * during the call (caller keeps ownership); output buffers are caller-allocated unless stated; no
 * callbacks, no exceptions across the boundary; calls on one context must be serialised by the caller
 * (same contract as the reference's non re-entrant callbacks, adpm.C:11-13).  There is NO CPU fallback:
 * rdc_create fails with RDC_E_NODEVICE when no CUDA device is usable.
 */
#ifndef RDC_H
#define RDC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rdc_ctx rdc_ctx;

/* ---- enumerations ---------------------------------------------------------------------------- */
enum rdc_model { RDC_ADPM = 0, RDC_PIHNA = 1, RDC_RIPF = 2, RDC_PROTEAS = 3, RDC_HCC = 4,
                 RDC_SOLID = 5 /* SolidSystem (solid_system.C): unknowns = current node positions x,y,z; see rdc_solid_* below */ };
enum rdc_elem  { RDC_TET4 = 4, RDC_HEX8 = 8 };
enum rdc_ksp   { RDC_KSP_GMRES = 0, RDC_KSP_CG = 1, RDC_KSP_BICGSTAB = 2 };
enum rdc_pc    { RDC_PC_JACOBI = 0, RDC_PC_NONE = 1, RDC_PC_BJACOBI = 2 /* reserved (v x v nodal block): RDC_E_ARG today */ };

enum rdc_status {
  RDC_OK = 0,
  RDC_E_ARG = -1,        /* invalid argument                                  */
  RDC_E_NODEVICE = -2,   /* no usable CUDA device (there is no CPU fallback)  */
  RDC_E_CUDA = -3,       /* CUDA runtime error, see rdc_last_error            */
  RDC_E_STATE = -4,      /* call sequence error (e.g. solve before assemble)  */
  RDC_E_NOMEM = -5,
  RDC_E_MESH = -6,       /* inconsistent mesh / dof map                       */
  RDC_E_DIVERGED = -7,   /* Krylov breakdown / divergence                     */
  RDC_E_COMM = -8,       /* NCCL failure                                      */
  RDC_E_MODEL = -9       /* model precondition violated (ripf.C:773: RT_total_max <= 0) */
};

/* ---- flat parameter vectors (rdc_set_params) --------------------------------------------------
 * One double per es.parameters value read by the reference callback, in the order of the reads. */

/* ADPM: adpm.C:368-414.  3-vectors are {cM, c0, c1} for Pi_/SD_ (utils.h:100-133), 5-vectors are
 * {cM, c0, c1, c2, c3} for Tr_ (utils.h:158-187).  Angles in radians (adpm.C:192,212 convert). */
enum {
  ADPM_GAMMA = 0,            /* decay/PrP/time_exponent                    adpm.C:368 */
  ADPM_DECAY_PRP = 1,        /* [3]  decay/PrP (before * pow(time,gamma))  adpm.C:369-371 */
  ADPM_DIFFUSE_AB = 4,       /* [3]  adpm.C:372-374 */
  ADPM_TAXIS1_AB = 7,        /* [3]  adpm.C:375-377 */
  ADPM_TAXIS2_AB = 10,       /* [3]  adpm.C:378-380 */
  ADPM_PRODUCE_AB = 13,      /* [3]  adpm.C:381-383 */
  ADPM_TRANSFORM_AB = 16,    /* [5]  adpm.C:384-388 */
  ADPM_DECAY_AB = 21,        /* [3]  adpm.C:389-391 */
  ADPM_DIFFUSE_TAU = 24,     /* [3]  adpm.C:392-394 */
  ADPM_TAXIS1_TAU = 27,      /* [3]  adpm.C:395-397 */
  ADPM_TAXIS2_TAU = 30,      /* [3]  adpm.C:398-400 */
  ADPM_PRODUCE_TAU = 33,     /* [3]  adpm.C:401-403 */
  ADPM_TRANSFORM_TAU = 36,   /* [5]  adpm.C:404-408 */
  ADPM_DECAY_TAU = 41,       /* [3]  adpm.C:409-411 */
  ADPM_ANGLE_AB = 44,        /* taxis/A_b/angle [rad]  adpm.C:413 */
  ADPM_ANGLE_TAU = 45,       /* taxis/Tau/angle [rad]  adpm.C:414 */
  ADPM_NPARAMS = 46
};

/* PIHNA: pihna.C:360-381 (necrosis rates are passed RAW; the / Kappa_k of pihna.C:364-366 is applied
 * by the library exactly as the callback does). */
enum {
  PIHNA_LAMBDA_K = 0,   /* cells_min_capacity */
  PIHNA_KAPPA_K = 1,    /* cells_max_capacity */
  PIHNA_KAPPA_A = 2,    /* cytokines_max_capacity */
  PIHNA_EK = 3,         /* cells_max_capacity/exponent */
  PIHNA_NECROSIS_C = 4, PIHNA_NECROSIS_H = 5, PIHNA_NECROSIS_V = 6,
  PIHNA_DIFFUSE_C = 7, PIHNA_TAXIS_C = 8, PIHNA_DIFFUSE_H = 9, PIHNA_TAXIS_H = 10,
  PIHNA_PRODUCE_C = 11, PIHNA_SWITCH_C2H = 12, PIHNA_SWITCH_H2C = 13, PIHNA_SWITCH_H2N = 14,
  PIHNA_DIFFUSE_V = 15, PIHNA_TAXIS_V = 16, PIHNA_PRODUCE_V = 17,
  PIHNA_SECRETE_A_C = 18, PIHNA_SECRETE_A_H = 19, PIHNA_UPTAKE_A_V = 20, PIHNA_DECAY_A = 21,
  PIHNA_NPARAMS = 22
};

/* RIPF: ripf.C:379-408 (assembly) + ripf.C:699-702 (check_solution).  fb/lambda/RT/r and
 * fb/omicro/RT/r are passed RAW: when 0 the library substitutes int(max RT_total) exactly as
 * ripf.C:398-403,772 do. */
enum {
  RIPF_VF_STROMA = 0, RIPF_VF_PARENCHYMA = 1, RIPF_VF_EXPONENT = 2, RIPF_VF_MIN_VACANT = 3,
  RIPF_VF_MAX_VACANT = 4, /* read but unused (ripf.C:383) */
  RIPF_PHI_CC_B = 5, RIPF_PHI_CC_D = 6, RIPF_PHI_CC = 7,
  RIPF_PHI_FB_B = 8, RIPF_PHI_FB_D = 9, RIPF_PHI_FB = 10, RIPF_PHI_TOL = 11,
  RIPF_KAPPA = 12, RIPF_KAPPA_RT_C = 13, RIPF_DELTA = 14, RIPF_DELTA_RT_A = 15, RIPF_DELTA_RT_B = 16,
  RIPF_LAMBDA = 17, RIPF_LAMBDA_RT_R = 18, RIPF_LAMBDA_HU_R = 19,
  RIPF_OMICRO = 20, RIPF_OMICRO_RT_R = 21, RIPF_OMICRO_FB_B = 22,
  RIPF_OMEGA = 23, RIPF_DIFFUSION = 24, RIPF_HAPTOTAXIS = 25, RIPF_RADIOTAXIS = 26,
  RIPF_HU_MIN = 27, RIPF_HU_MAX = 28,                       /* ripf.C:699-700 */
  RIPF_RT_BROAD_FRAC = 29, RIPF_RT_FOCUS_FRAC = 30,         /* ripf.C:701-702 (ints stored as double) */
  RIPF_NPARAMS = 31
};

/* PROTEAS: proteas.C:378-410 */
enum {
  PROTEAS_T_MAX = 0, PROTEAS_RT_MAX = 1,
  PROTEAS_RHO_H = 2, PROTEAS_U_H = 3, PROTEAS_DELTA_H = 4, PROTEAS_A_RT_H = 5, PROTEAS_B_RT_H = 6,
  PROTEAS_NU_H = 7,
  PROTEAS_D_C = 8, PROTEAS_D_C_H = 9, PROTEAS_RHO_C = 10, PROTEAS_U_C = 11, PROTEAS_DELTA_C = 12,
  PROTEAS_A_RT_C = 13, PROTEAS_B_RT_C = 14, PROTEAS_NU_C = 15,
  PROTEAS_PSI_N = 16, PROTEAS_K_N = 17, PROTEAS_U_N = 18,
  PROTEAS_RHO_V = 19, PROTEAS_NU_V = 20,
  PROTEAS_D_E = 21, PROTEAS_RHO_E = 22, PROTEAS_U_E = 23, PROTEAS_XI_E = 24, PROTEAS_P_RT_E = 25,
  PROTEAS_PSI_E = 26,
  PROTEAS_NPARAMS = 27
};

/* HCC: coupled_hcc.C:452-461 (necrosis rates RAW; / Kappa_k applied by the library, :459-461) */
enum {
  HCC_LAMBDA_K = 0, HCC_KAPPA_K = 1, HCC_EK = 2, HCC_PRODUCE_L = 3,
  HCC_DIFFUSE_C = 4, HCC_MECHANO_C = 5, HCC_PRODUCE_C = 6,
  HCC_NECROSIS_L = 7, HCC_NECROSIS_C = 8, HCC_NECROSIS_P = 9,
  HCC_NPARAMS = 10
};

/* number of unknowns per node / parameter count for a model */
int rdc_model_nvars(int model);     /* 3 (ADPM, RIPF, HCC) or 5 (PIHNA, PROTEAS) */
int rdc_model_nparams(int model);

/* ---- life cycle ------------------------------------------------------------------------------ */

/* Flattened hand-over (done once).  conn uses the libMesh local node order (identical to Gmsh's for
 * TET4/HEX8).  node_dof_base[n] is the global dof id of variable 0 at node n; variables of a node are
 * contiguous (dof = base + var, libMesh variable groups) -- NULL means base = nvars * n.
 * device < 0 selects the current CUDA device.  The context owns all device memory. */
int rdc_create(rdc_ctx** out, int model, int elem_type,
               int64_t n_nodes, int64_t n_elems,
               const int32_t* conn,           /* [n_elems * nen]            */
               const double* xyz,             /* [n_nodes * 3]              */
               const int32_t* node_dof_base,  /* [n_nodes] or NULL          */
               int device);

/* Same, for rank `rank` of an `nranks`-process job (one process per GPU).  Every rank passes the SAME
 * full flattened mesh (the reference keeps a replicated Mesh, adpm.C:17); the library partitions the
 * nodes (METIS k-way on the nodal graph, or recursive coordinate bisection when partitioner = 1),
 * keeps the rows of its owned nodes plus the ghost layer, and exchanges ghost values over NCCL.
 * nccl_unique_id is the 128-byte ncclUniqueId made by rdc_comm_unique_id on rank 0 and distributed by
 * the caller (MPI_Bcast in rdcFEs, torch.distributed in bench.py). */
int rdc_create_distributed(rdc_ctx** out, int model, int elem_type,
               int64_t n_nodes, int64_t n_elems,
               const int32_t* conn, const double* xyz, const int32_t* node_dof_base,
               int device, int rank, int nranks, int partitioner,
               const void* nccl_unique_id /* 128 bytes */);
int rdc_comm_unique_id(void* out128);

void rdc_destroy(rdc_ctx*);
const char* rdc_last_error(const rdc_ctx*);    /* ctx may be NULL: message of the last failed create */

/* ---- data hand-over --------------------------------------------------------------------------- */
int rdc_set_params(rdc_ctx*, const double* p, int n);
/* slot 0: ADPM tract vectors, ncomp = 3, element order = conn order */
int rdc_set_elem_field(rdc_ctx*, int slot, const double* f /* [n_elems*ncomp] */, int ncomp);
/* slot 0: RIPF RT dose, ncomp = 2 {broad, focus} (ripf.C:275-289); PROTEAS AUX, ncomp = 2 {HU, RTD} */
int rdc_set_nodal_field(rdc_ctx*, int slot, const double* f /* [n_nodes*ncomp] */, int ncomp);
int rdc_update_coords(rdc_ctx*, const double* xyz /* [n_nodes*3] */);
/* u is indexed by global dof id (length rdc_n_dofs).  Sets solution == current_local_solution. */
int rdc_set_solution(rdc_ctx*, const double* u);
int rdc_get_solution(rdc_ctx*, double* u);      /* distributed: every rank receives the full vector */
/* distributed: writes only the entries of the dofs owned by this rank (no all-gather); single rank: == rdc_get_solution.
 * Replaces the rank-local part of system.solution that libMesh keeps per processor (PETSc MPI vector). */
int rdc_get_solution_owned(rdc_ctx*, double* u);
int rdc_get_old_solution(rdc_ctx*, double* u);
int64_t rdc_n_dofs(const rdc_ctx*);
int rdc_set_time(rdc_ctx*, double time);        /* system.time (adpm.C:64) */
int rdc_set_dt(rdc_ctx*, double dt);            /* es.parameters "time_step"; needed before RIPF's pre-loop rdc_clamp (ripf.C:53,697) */

/* ---- the hot path ----------------------------------------------------------------------------- */
int rdc_rotate(rdc_ctx*);                                   /* older <- old <- current */
int rdc_assemble(rdc_ctx*, double time, double dt);         /* K, F from old solution; device resident */
/* KSPSolve: initial guess = current solution; converged when ||B r|| <= max(rtol ||B b||, 1e-50) (PETSc default test,
 * left preconditioning).  `restart` is used by GMRES only (<= 0: 30).  A BiCGStab breakdown (rho or omega = 0)
 * continues with GMRES from the current iterate; RDC_E_DIVERGED is returned only for a NaN residual. */
int rdc_solve(rdc_ctx*, int ksp, int pc, double rtol, int maxits, int restart,
              int* iterations, double* resnorm);
int rdc_clamp(rdc_ctx*);                                    /* the model's check_solution, at ctx time */
/* time += dt is the caller's business (adpm.C:63): pass the NEW time. rotate+assemble+solve+clamp. */
int rdc_step(rdc_ctx*, double time, double dt, int ksp, int pc, double rtol, int maxits, int restart,
             int* iterations, double* resnorm);
/* y = K x with the assembled operator (global dof order, host buffers) -- parity / roofline probe */
int rdc_spmv(rdc_ctx*, const double* x, double* y);
/* device-resident repetition of the SpMV kernel for the roofline measurement: returns mean ms */
int rdc_bench_spmv(rdc_ctx*, int reps, double* mean_ms);

/* ---- post-step reductions of save_solution (adpm.C:690-829, pihna.C:842-976, ripf.C:777-864) --------------------
 * Device-side replacement of the rank-0 serial element loops behind the per-case CSV output; results are the same on
 * every rank.  rdc_set_subdomains hands over elem->subdomain_id() renumbered 0..n_regions-1 (NULL: one region; it is
 * the default).  A condition holds at a node when  lo <= (sum_a w[a]*u[a]) / div <= hi  (non-zero weights, ascending
 * variable order -- c+h, (n+c+h+v)/Kappa_k, HU, cc >= min ...); an element counts when every condition holds at every
 * one of its nodes.  rdc_region_last_mean reproduces adpm.C:780-783: the element average of variable `var` in the
 * LAST element (element id order) of every region. */
struct rdc_range_cond { double w[5]; double div, lo, hi; };
int rdc_set_subdomains(rdc_ctx*, const int32_t* region /* [n_elems] or NULL */, int n_regions);
int rdc_region_volumes(rdc_ctx*, int ncond, const struct rdc_range_cond* cond, double* vol /* [n_regions] */);
int rdc_region_last_mean(rdc_ctx*, int var, double* mean /* [n_regions] */);

/* ---- solid mechanics: SolidSystem of solid_system.C / solid.C / coupled_hcc.C:117-132 -----------
 * A context created with model RDC_SOLID holds the CURRENT node positions as its solution (rdc_set_solution /
 * rdc_get_solution, dof = 3*node + d unless node_dof_base says otherwise) -- solid.C:27-30, mesh_position_get/set.  Works
 * with rdc_create_distributed too (one process per GPU: owned rows, ghost positions and step vectors exchanged like the RDC
 * systems', norms all-reduced in rank order so that every rank takes the same Newton decisions).  The Krylov solvers of
 * rdc_solve (ksp) with point Jacobi solve the Newton systems.
 *   rdc_solid_set_reference  <- SolidSystem::save_initial_mesh (solid_system.C:26-48): undeformed positions [n_nodes*3]
 *   rdc_solid_set_materials  <- es.parameters "material/<id>/Hyperelastic/{Young,Poisson,FibreStiffness,
 *                               VolumetricStretchRatio/rate_0..2}" (solid.C:276-291, read at solid_system.C:182-189):
 *                               mats[nmat*6] in that order, mat_of[n_elems] = index of elem->subdomain_id() in the table
 *   rdc_solid_set_fibres     <- "SolidSystem::fibre" variables 0-2 (solid.C:303-337, solid_system.C:205-213): [n_elems*3] or NULL
 *   rdc_solid_set_bcs        <- "BCs", "BC/<id>/displacement", "BCs/displacement_penalty" + BoundaryInfo::has_boundary_id
 *                               (solid_system.C:288-304): side k = libMesh side side_no[k] of element side_elem[k] carries
 *                               condition side_bc[k]; bc_disp[nbc*3], NaN = component left free
 *   rdc_solid_assemble       <- [upstream] FEMSystem::assembly(true, true): element_time_derivative + side_time_derivative
 *                               (solid_system.C:146-371) -> Jacobian (rdc_download_csr) and residual (rdc_get_rhs)
 *   rdc_solid_newton         <- SolidSystem::run_solver (solid_system.C:373-392) = [upstream] NewtonSolver::solve configured by
 *                               solid_system.C:80-98; opts = {max_nonlinear_iterations, relative_step_tolerance,
 *                               relative_residual_tolerance, absolute_residual_tolerance, require_reduction,
 *                               max_linear_iterations, initial_linear_tolerance}; info[4] = {newton iterations, linear
 *                               iterations, final residual norm, converged 0/1}
 *   rdc_solid_post_process   <- SolidSystem::post_process (solid_system.C:394-538): per element mean normal stress
 *                               ("SolidSystem::pressure"), von Mises stress, current fibre vector; outputs may be NULL */
int rdc_solid_set_reference(rdc_ctx*, const double* xyz_undeformed /* [n_nodes*3] */);
int rdc_solid_set_materials(rdc_ctx*, int nmat, const double* mats /* [nmat*6] */, const int32_t* mat_of /* [n_elems] or NULL */);
int rdc_solid_set_fibres(rdc_ctx*, const double* fibres /* [n_elems*3] or NULL */);
/* es.parameters "solver/assembly_use_symmetry" (solid.C:243-244; solid_system.C:180,248-262): 1 = the blocks j >= i of every
 * element tangent are evaluated and the others mirrored (K_ji = K_ij^T) like the reference does; default 0 */
int rdc_solid_set_symmetry(rdc_ctx*, int use_symmetry);
int rdc_solid_set_bcs(rdc_ctx*, int nbc, const double* bc_disp /* [nbc*3] */, int64_t nside, const int64_t* side_elem,
                      const int32_t* side_no, const int32_t* side_bc, double penalty);
int rdc_solid_assemble(rdc_ctx*, double pseudo_time);
int rdc_solid_newton(rdc_ctx*, double pseudo_time, const double* opts /* [7] */, int ksp, double* info /* [4] */);
int rdc_solid_post_process(rdc_ctx*, double pseudo_time, double* press, double* von_mises, double* fibre_current);
/* host-only probes of the element arithmetic (no device needed; the CPU tests hold them to the oracle): row li of the element
 * residual/tangent R[3], K[9*nen] (plane a*3+c, column node j at (a*3+c)*nen + j); the penalty row of node i of a side with
 * ns nodes R[3], Kd[ns*3]; post_process of one element out[5] = {p, von Mises, fibre[3]} */
int rdc_solid_probe_row(int elem_type, const double* x_cur, const double* x_und, const double* mat6, double pseudo_time,
                        const double* eta, int li, int use_symmetry, double* R, double* K);
int rdc_solid_probe_bc_row(int ns, const double* x_cur, const double* x_und, const double* disp, double pseudo_time, double penalty,
                           int i, double* R, double* Kd);
int rdc_solid_probe_post(int elem_type, const double* x_cur, const double* x_und, const double* mat6, double pseudo_time,
                         const double* eta, double* out5);
/* host-only probe of the penalty-row lists of rank `rank` of an nranks job (CPU world-size-2 tests): rows as global node ids,
 * entries (side index into the input, position of the node inside that side) in the order the terms are added; rdc_free */
int rdc_solid_probe_bc_rows(int elem_type, int64_t n_nodes, int64_t n_elems, const int32_t* conn, const double* xyz, int rank,
                            int nranks, int partitioner, int64_t nside, const int64_t* side_elem, const int32_t* side_no,
                            int32_t* n_rows, int32_t** row_node_glob, int32_t** row_ptr, int32_t** ent_side, int32_t** ent_pos);

/* read-only streaming probe over the stored operator values (reference point for the SpMV roofline): mean ms, bytes */
int rdc_bench_stream(rdc_ctx*, int reps, int ctas_per_sm, double* mean_ms, int64_t* bytes);

/* measured rate of the fp64 pipe (independent DFMA chains, full occupancy), TFLOP/s: the compute roofline of the assembly kernel */
int rdc_bench_dfma(rdc_ctx*, double* tflops);
/* cost of one grid barrier (mode 0) / reduction-barrier (mode 1) of the persistent Krylov kernel on this device: mean us */
int rdc_bench_barrier(rdc_ctx*, int reps, int ctas_per_sm, int mode, double* mean_us);

/* ---- parity / introspection ------------------------------------------------------------------- */
/* Scalar CSR in global dof numbering, rows and columns sorted: exactly the (node graph + I) (x) dense
 * v x v pattern libMesh preallocates.  Buffers are malloc'ed by the library; free
 * with rdc_free.  Distributed: only the rows owned by this rank (global ids in rows[]). */
int rdc_download_csr(rdc_ctx*, int64_t* n_rows, int64_t* nnz,
                     int64_t** rows, int64_t** rowptr, int32_t** col, double** val, double** rhs);
void rdc_free(void*);
/* the assembled load vector F in global dof numbering (length rdc_n_dofs; distributed: every rank receives all of it):
 * system.rhs after the assemble callback -- lets a caller form the true residual F - K u with rdc_spmv */
int rdc_get_rhs(rdc_ctx*, double* rhs);

struct rdc_stats {
  double ms_assemble;       /* last rdc_assemble, CUDA events on the context stream   */
  double ms_solve;          /* last rdc_solve                                          */
  double ms_clamp;
  double ms_spmv_total;     /* sum over the SpMV launches of the last solve            */
  int    n_spmv;            /* SpMV launches in the last solve                         */
  int    iterations;        /* Krylov iterations of the last solve                     */
  double resnorm;           /* final (preconditioned) residual norm                    */
  double resnorm0;          /* reference norm ||B b|| used by the convergence test     */
  int64_t n_nodes_local, n_nodes_ghost, n_elems_local, nnzb_local;
  int64_t bytes_assemble;   /* algorithmic bytes of one assembly    */
  int64_t bytes_spmv;       /* algorithmic bytes of one SpMV in the traversed format   */
  int64_t bytes_index;      /* index overhead read per assembly (maps), reported separately */
  int64_t kernel_launches;  /* kernels launched by this context so far                 */
  int     ripf_rt_total_max;/* RIPF: int(max RT_total) (ripf.C:772)                    */
  /* running totals since rdc_create (phase times are resolved lazily, without host synchronisation inside a step) */
  double sum_ms_assemble, sum_ms_solve, sum_ms_clamp, sum_ms_spmv;
  int64_t sum_iterations, sum_n_spmv, n_solves;
  /* distributed runs: 1 when ghost values and dot products travel over NVLink peer memory (cudaIpc arenas), 0 on the
   * NCCL transport; p2p_fused = the exchanges are finished inside the producing Krylov kernels */
  int     p2p_on, p2p_fused;
  int     bicg_persistent;  /* the last BiCGStab solve ran as one cooperative launch (its SpMV time is then measured with %globaltimer in the kernel) */
};
int rdc_get_stats(rdc_ctx*, struct rdc_stats*);
/* Host-only probe of the node partition and halo lists of rank `rank` (no device needed; used by the CPU
 * world_size-2 tests).  All arrays are malloc'ed by the library (rdc_free).  send_glob/recv_glob are global node
 * ids grouped per neighbour by send_ptr/recv_ptr ([n_nbr+1] offsets). */
int rdc_probe_partition(int elem_type, int nvars, int64_t n_nodes, int64_t n_elems, const int32_t* conn,
                        const double* xyz, int rank, int nranks, int partitioner, int32_t* n_owned, int32_t* n_ghost,
                        int64_t* n_elems_local, int32_t** owner, int32_t* n_nbr, int32_t** nbr_rank,
                        int32_t** send_ptr, int32_t** send_glob, int32_t** recv_ptr, int32_t** recv_glob);
/* Tuning switch of the context (defaults: environment RDC_<NAME>): "spmv_tma" 1/0 (TMA-staged or LDG SpMV),
 * "spmv_minb", "spmv_ctas_per_sm", "tma_ctas_per_sm", "tma_stages", "sync_every", "p2p_fused_ar",
 * "p2p_fused_halo", "bicg_persist" 1/0/-1 (BiCGStab as one cooperative launch with grid barriers, five launches per iteration, or
 * chosen by the problem size per GPU), "node_order" 1/0 (read at rdc_create from RDC_NODE_ORDER: owned nodes numbered along a
 * Morton curve or by ascending global id), "trace".  Results do not depend on them beyond floating-point summation order. */
int rdc_set_option(rdc_ctx*, const char* name, int value);
/* Host-only probes of set-up logic, for CPU tests (arrays malloc'ed by the library, rdc_free): the SpMV tile cutter
 * (tiles = n_tiles x {row0, nrows, first block, nblocks}; n_tiles = -1 when a row exceeds max_blocks) and the region
 * bucketing of the save_solution reductions. */
int rdc_probe_spmv_tiles(int32_t n_rows, const int32_t* rowptr, int max_rows, int max_blocks, int32_t* n_tiles, int32_t** tiles);
int rdc_probe_region_chunks(int64_t n_elems, const uint8_t* counted, const int32_t* region, int n_regions, int chunk,
                            int64_t* n_counted, int32_t** perm, int32_t* n_chunks, int32_t** chunk_ptr, int32_t** rchunk_ptr);
/* run every kernel on this cudaStream_t (default: a stream owned by the context) */
int rdc_set_stream(rdc_ctx*, void* cuda_stream);
const char* rdc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RDC_H */
