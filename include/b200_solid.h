/*
 * b200_solid.h -- C ABI of lib4c_b200_solid.so
 *
 * B200-native (sm_100a, FP64) replacement for ONE hot path of 4C: the per-Newton-iteration
 * evaluation of solid element internal force + tangent stiffness and its owner-computes
 * scatter-add into the CSR values of Core::LinAlg::SparseMatrix and the residual vector.
 *
 * Every entry point names the reference interface it replaces (paths relative to the 4C tree,
 * commit 96b9fa91).  Plain pointers and sizes only; no exceptions cross the boundary: every call
 * returns B200S_OK (0) or a negative b200s_status and b200s_last_error() holds the text.  One
 * context <-> one rank <-> one GPU; calls on a context are made from one host thread (the
 * reference element loop is single threaded per rank, 4C_mat_elasthyper_service.cpp:31-36).
 *
 * Pointers marked [host|dev] may live in host or device memory (resolved by UVA).
 */
#ifndef B200_SOLID_H
#define B200_SOLID_H

#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

  typedef struct b200s_ctx b200s_ctx;

  typedef enum
  {
    B200S_OK = 0,
    B200S_ERR_INVALID = -1,        /* bad argument / call order */
    B200S_ERR_CUDA = -2,           /* CUDA runtime failure (text in b200s_last_error) */
    B200S_ERR_UNSUPPORTED = -3,    /* cell/material/graph layout outside the device path:
                                      the adapter must fall back to the CPU element loop */
    B200S_ERR_NEG_JACOBIAN = -4,   /* FOUR_C_ASSERT_ALWAYS(det > 0), 4C_solid_ele_calc_lib.hpp:558 */
    B200S_ERR_NONFINITE = -5,      /* NaN/Inf in f or K (FE trapping, model_evaluator_generic.cpp:229) */
    B200S_ERR_COMM = -6            /* halo exchange failure */
  } b200s_status;

  /* Core::FE::CellType values on the device path (4C_fem_general_cell_type.hpp) */
  typedef enum
  {
    B200S_HEX8 = 0,
    B200S_HEX27 = 1,
    B200S_TET10 = 2,
    /* displacement-based blocks of further cell types (create_solid_calculation_interface,
     * 4C_solid_ele_factory.hpp:36-38,79-89) run through the table-driven general kernel (no specialised
     * fast kernel; KINEM nonlinear / linear, every material kind, no element technology / prestress): */
    B200S_TET4 = 3, /* tet_1point stiffness rule, tet_4point mass rule */
    B200S_HEX20 = 4, /* hex_27point */
    B200S_WEDGE6 = 5,  /* wedge_6point; general kernel, no surface loads (two surface shapes) */
    B200S_PYRAMID5 = 6, /* pyramid_8point, the reference's rational shape functions; general kernel, no surface loads */
    B200S_HEX18 = 7     /* quad9 in the plane x linear through the thickness, hex_18point; general kernel, no surface loads */
  } b200s_celltype;

  /* Mat::So3Material implementations on the device path (Core::Materials::MaterialType):
   *   COUP_NEOHOOKE : MAT_ElastHyper{ELAST_CoupNeoHooke YOUNG NUE}     params = {YOUNG, NUE}
   *   ISO_VOL       : MAT_ElastHyper{ELAST_IsoNeoHooke MUE, ELAST_VolSussmanBathe KAPPA}
   *                   params = {MUE, KAPPA}; MUE may be overridden per element (InputField)
   *   SVK           : MAT_Struct_StVenantKirchhoff YOUNG NUE           params = {YOUNG, NUE}
   *   ELASTHYPER    : MAT_ElastHyper with a run-time list of isotropic summands, installed with
   *                   b200s_set_material_elasthyper (slower than the three specialised kinds)
   *   DRUCKER_PRAGER: MAT_Struct_DruckerPrager (src/mat/4C_mat_plasticdruckerprager.cpp), a HISTORY material:
   *                   params = {YOUNG, NUE, ISOHARD, TOL, C, ETA, XI, ETABAR, MAXITER, TANG (0 elastic,
   *                   1 consistent)}; small strains, i.e. KINEM linear blocks only.  Plastic strain and
   *                   accumulated plastic strain of every Gauss point (last converged + current iterate) live
   *                   in the context: b200s_update_step = Mat::PlasticDruckerPrager::update(),
   *                   b200s_reset_step discards the iterate, b200s_get/set_material_history = pack/unpack.
   *                   With ETA != ETABAR the consistent tangent is not major-symmetric: such blocks are evaluated
   *                   by the general kernel (full 6x6 tangent, full K_e; slower than the symmetric fast path).
   *   VISCONEOHOOKE : MAT_VISCONEOHOOKE (src/mat/4C_mat_visconeohooke.cpp), a finite-strain HISTORY material:
   *                   params = {YOUNGS_SLOW, POISSON, DENS, YOUNGS_FAST, RELAX, THETA}; compressible Neo-Hooke on
   *                   the fast modulus whose deviatoric stress relaxes through one Maxwell branch (one-step-theta
   *                   rule over the time step given with b200s_set_time_step_size).  KINEM nonlinear blocks, TECH
   *                   none (hex8 / hex27 / tet10) or eas_mild / eas_full (hex8).  24 doubles of state per Gauss
   *                   point (last converged and current {S_dev, Q}); b200s_update_step =
   *                   Mat::ViscoNeoHooke::update(), b200s_get/set_material_history = pack/unpack.
   *   PLASTIC_NLNLOG: MAT_Struct_PlasticNlnLogNeoHooke (src/mat/4C_mat_plasticnlnlogneohooke.cpp), finite-strain von
   *                   Mises plasticity in principal logarithmic strains, a HISTORY material: params = {YOUNG, NUE,
   *                   DENS, YIELD, ISOHARD, SATHARDENING, HARDEXPO, VISC, RATE_DEPENDENCY, TOL} (HARDENING_FUNC, a
   *                   parsed function: CPU path); the rate term uses the time step of b200s_set_time_step_size.
   *                   KINEM nonlinear blocks, TECH none (every cell; hex8 on the fast kernel, the others on the
   *                   general kernel) or eas_mild / eas_full (hex8).  14 doubles of state per Gauss point in the
   *                   reference's pack order {accplstrainlast, invplrcglast[6], accplstraincurr, invplrcgcurr[6]}
   *                   (tensor components 00,11,22,01,12,02; a fresh point holds the identity); b200s_update_step =
   *                   update(), b200s_get/set_material_history = pack/unpack. */
  typedef enum
  {
    B200S_MAT_COUP_NEOHOOKE = 0,
    B200S_MAT_ISO_VOL = 1,
    B200S_MAT_SVK = 2,
    B200S_MAT_ELASTHYPER = 3,
    B200S_MAT_DRUCKER_PRAGER = 4,
    B200S_MAT_VISCONEOHOOKE = 5,
    B200S_MAT_PLASTIC_NLNLOG = 6
  } b200s_material;

  /* Mat::Elastic::Summand implementations evaluated through the principal / modified invariants
   * (src/mat/elast/4C_mat_elast_*.cpp; add_derivatives_principal / add_derivatives_modified /
   * add_strain_energy).  Parameters in input-file order, at most 4 per summand:
   *   COUP_NEOHOOKE {YOUNG,NUE}  COUP_MOONEYRIVLIN {C1,C2,C3}  COUP_LOGNEOHOOKE {C1,C2} + MODE
   *   COUP_LOGMIXNEOHOOKE {C1,C2} + MODE  COUP_SVK {YOUNG,NUE}  COUP_SIMOPISTER {MUE}
   *   COUP_BLATZKO {MUE,NUE,F}  COUP_EXPPOL {A,B,C}  COUP_1POW/2POW/3POW {C,D}  COUP_13APOW {C,D,A}
   *   ISO_NEOHOOKE {MUE}  ISO_MOONEYRIVLIN {C1,C2}  ISO_YEOH {C1,C2,C3}  ISO_1POW/2POW {C,D}
   *   ISO_EXPOPOW {K1,K2,C}  VOL_SUSSMANBATHE {KAPPA}  VOL_PENALTY {EPSILON,GAMMA}
   *   VOL_OGDEN {KAPPA,BETA}  VOL_POW {A,EXPON}
   * MODE (LogNeoHooke / LogMixNeoHooke): 0 = "YN" (C1 = Young, C2 = nue), 1 = "Lame" (C1 = mue,
   * C2 = lambda).  Stretch-based summands (IsoOgden, IsoVarga, CoupVarga) and fibre summands other than
   * COUP_ANISOEXPO: CPU path. */
  typedef enum
  {
    B200S_SUM_COUP_NEOHOOKE = 0,
    B200S_SUM_COUP_MOONEYRIVLIN = 1,
    B200S_SUM_COUP_LOGNEOHOOKE = 2,
    B200S_SUM_COUP_LOGMIXNEOHOOKE = 3,
    B200S_SUM_COUP_SVK = 4,
    B200S_SUM_COUP_SIMOPISTER = 5,
    B200S_SUM_COUP_BLATZKO = 6,
    B200S_SUM_COUP_EXPPOL = 7,
    B200S_SUM_COUP_1POW = 8,
    B200S_SUM_COUP_2POW = 9,
    B200S_SUM_COUP_3POW = 10,
    B200S_SUM_COUP_13APOW = 11,
    B200S_SUM_ISO_NEOHOOKE = 12,
    B200S_SUM_ISO_MOONEYRIVLIN = 13,
    B200S_SUM_ISO_YEOH = 14,
    B200S_SUM_ISO_1POW = 15,
    B200S_SUM_ISO_2POW = 16,
    B200S_SUM_ISO_EXPOPOW = 17,
    B200S_SUM_VOL_SUSSMANBATHE = 18,
    B200S_SUM_VOL_PENALTY = 19,
    B200S_SUM_VOL_OGDEN = 20,
    B200S_SUM_VOL_POW = 21,
    B200S_SUM_COUP_ANISOEXPO = 22, /* {K1,K2,K1COMP,K2COMP}: ELAST_CoupAnisoExpo on the element fibre FIBER1
                                      (INIT 1, standard structural tensor strategy; needs
                                      b200s_set_element_fibers) */
    /* summands in (modified) principal stretches: Mat::elast_hyper_add_response_stretches
     * (src/mat/4C_mat_elasthyper_service.cpp:208-386).  Like in the reference they have no strain energy:
     * b200s_calc_energy returns B200S_ERR_UNSUPPORTED ("Summand does not support calculation of strain
     * energy", src/mat/elast/4C_mat_elast_summand.hpp:127-139). */
    B200S_SUM_ISO_OGDEN = 23,  /* {MUE, ALPHA}  ELAST_IsoOgden  */
    B200S_SUM_COUP_VARGA = 24, /* {MUE, BETA}   ELAST_CoupVarga */
    B200S_SUM_ISO_VARGA = 25   /* {MUE, BETA}   ELAST_IsoVarga  */
  } b200s_summand;

  /* Discret::Elements::SolidElementProperties of an element block (4C_solid_ele_properties.hpp:30-65;
   * input keys KINEM / TECH / PRESTRESS_TECH of the SOLID element line) */
  typedef enum
  {
    B200S_KINEM_NONLINEAR_TOTLAG = 0, /* KINEM nonlinear */
    B200S_KINEM_LINEAR = 1            /* KINEM linear: displacement based (4C_solid_ele_calc_displacement_based_
                                         linear_kinematics.hpp) or hex8 TECH eas_mild / eas_full (EASFormulation<...,
                                         linear>, 4C_solid_ele_calc_eas_helpers.hpp:648-666) */
  } b200s_kinem_type;
  typedef enum
  {
    B200S_TECH_NONE = 0,
    B200S_TECH_FBAR = 1,     /* hex8, pyramid5: 4C_solid_ele_calc_fbar.hpp (pyramid5 and the MULF combination: general kernel) */
    B200S_TECH_EAS_MILD = 2, /* hex8, 9 enhanced strain parameters: 4C_solid_ele_calc_eas.hpp */
    B200S_TECH_EAS_FULL = 3, /* hex8, 21 parameters */
    B200S_TECH_SHELL_ANS = 4, /* hex8 solid shell with assumed natural strains (thickness = t-direction of the element):
                                 4C_solid_ele_calc_shell_ans.hpp; general kernel.  wedge6 shell_ans, shell_eas_ans: CPU path */
    B200S_TECH_SHELL_EAS = 5, /* hex8, EAS with the 7 solid-shell modes (eastype_sh8_7, t = thickness direction):
                                 EASSolidIntegrator<hex8, eastype_sh8_7, nonlinearTotLag>; the FMA EAS kernel */
    B200S_TECH_SHELL_EAS_ANS = 6 /* hex8 solid shell with ANS and the 7 EAS modes condensed with the ANS B-operator:
                                    EasAnsSolidShellIntegrator<hex8, eastype_sh8_7> (4C_solid_ele_calc_shell_eas_ans.hpp);
                                    general kernel.  wedge6 (eastype_sw6_1): CPU path */
  } b200s_element_technology;
  typedef enum
  {
    B200S_PRESTRESS_NONE = 0,
    B200S_PRESTRESS_MULF = 1 /* 4C_solid_ele_calc_mulf.hpp (every displacement-based cell); with B200S_TECH_FBAR:
                                4C_solid_ele_calc_mulf_fbar.hpp (hex8, pyramid5) -- the Gauss-point history then carries one more
                                MulfHistoryData per element, the global history at the centroid, after its Gauss points */
  } b200s_prestress_technology;

  /* Core::Elements::ActionType values handled by b200s_evaluate (legacy_enum_definitions;
   * 4C_solid_ele_evaluate.cpp:64-159).  struct_calc_stress and struct_calc_energy have their own
   * entry points (b200s_calc_stress, b200s_calc_energy). */
  typedef enum
  {
    B200S_ACTION_INTERNALFORCE = 0,        /* struct_calc_internalforce */
    B200S_ACTION_NLNSTIFF = 1,             /* struct_calc_nlnstiff */
    B200S_ACTION_NLNSTIFFMASS = 2,         /* struct_calc_nlnstiffmass: + mass matrix, inertia force */
    B200S_ACTION_NLNSTIFFLMASS = 3,        /* struct_calc_nlnstifflmass: lumped mass */
    B200S_ACTION_INTERNALINERTIAFORCE = 4  /* struct_calc_internalinertiaforce: f_int and M_e * acc */
  } b200s_action;

  /* FourC::Solid::StressType / StrainType of struct_calc_stress (4C_solid_ele_calc_lib_io.hpp:140-290) */
  typedef enum
  {
    B200S_STRESS_NONE = 0,
    B200S_STRESS_2PK = 1,
    B200S_STRESS_CAUCHY = 2
  } b200s_stress_type;
  typedef enum
  {
    B200S_STRAIN_NONE = 0,
    B200S_STRAIN_GL = 1,
    B200S_STRAIN_EA = 2 /* strain_log (eigen decomposition): CPU path */
  } b200s_strain_type;

  /* which system matrix of the context an operation addresses */
  typedef enum
  {
    B200S_MATRIX_STIFFNESS = 0,
    B200S_MATRIX_MASS = 1
  } b200s_matrix;

  /* TYPE of a DESIGN SURF NEUMANN CONDITIONS entry (SolidSurface::evaluate_neumann,
   * src/solid_ele/4C_solid_ele_surface_evaluate.cpp:61-106). */
  typedef enum
  {
    B200S_NEUMANN_LIVE = 0,                 /* dead load on the material configuration */
    B200S_NEUMANN_PSEUDO_ORTHOPRESSURE = 1, /* pressure on the last converged configuration, no linearisation */
    B200S_NEUMANN_ORTHOPRESSURE = 2         /* follower pressure on the spatial configuration (LOADLIN) */
  } b200s_neumann_type;
  /* flags of b200s_evaluate_neumann */
#define B200S_NEUMANN_LOADLIN 1     /* elemat1 != nullptr: add the load stiffness into the stiffness matrix */
#define B200S_NEUMANN_MULF_ACTIVE 2 /* is_mulf_active(time): orthopressure acts on the material configuration */

  /* Mirror of the element error interface (4C_structure_new_elements_paramsinterface.hpp:46-55). */
  typedef struct
  {
    int32_t code;              /* b200s_status of the evaluate */
    int32_t n_neg_jacobian;    /* elements with det J <= 0 at an element node or Gauss point */
    int64_t first_bad_element; /* column-element index of one such element, -1 if none */
    double min_detJ;           /* over all nodes / Gauss points evaluated */
    double min_detF;
    int64_t n_nonfinite;       /* only filled when b200s_set_option(ctx,"check_finite",1) */
  } b200s_error_info;

  /* ---- life cycle -------------------------------------------------------------------------- */
  /* One context per rank/GPU.  rank/nranks describe the partition (owner-computes); with
   * nranks==1 no halo is needed. */
  int b200s_create(int device, int rank, int nranks, b200s_ctx** out);
  void b200s_destroy(b200s_ctx* ctx);
  const char* b200s_last_error(const b200s_ctx* ctx);
  /* options: "check_finite" (0/1: count NaN/Inf in f and K after each evaluate),
   * "reorder" (default 1: the device visits the elements in a locality (Morton) order chosen at the
   * next b200s_set_mesh; results, element indices in b200s_error_info and per-element parameters
   * stay in the caller's numbering), "double_buffer" (default 0 = one CSR value buffer, cleared at the
   * start of every evaluate; 1 = two buffers, the zero() of the next Newton iteration's matrix runs on a
   * side stream while the current evaluate fills the other buffer and b200s_device_values() alternates
   * between two addresses -- measured on B200: the concurrent memset slows the element kernel by as much
   * as it saves (2.46 vs 2.48 ms per step), hence off; takes effect at the next set_graph / build_graph),
   * "pipeline_phases" (default 12; at the next set_graph / build_graph the elements are grouped by the CSR
   * row range -- about nnz / phases entries each -- of their lowest owned node, so that b200s_evaluate with a
   * HOST K_values copies the rows of a finished phase while the next phases are still being evaluated;
   * 1 = one phase, no overlap; meshes whose row numbering has no spatial locality get one phase),
   * "symmetric_transfer" (default 0; with a HOST K_values and a formulation whose K_e is symmetric -- everything on
   * the device path except TECH fbar -- only the 3x3 blocks with column node >= row node cross PCIe, about half
   * of the bytes, and a pool of "host_threads" threads (default: all, at most 32) writes the full CSR rows into
   * K_values: the upper part of a row is copied, a lower block is the transpose of an upper block of an earlier
   * row; "nt_stores" (default 1) uses non-temporal stores for that.  Needs a node-blocked, structurally symmetric
   * graph whose owned column ids equal the row ids; otherwise the full transfer is used,
   * b200s_symmetric_transfer_status says why),
   * "overlap_halo" (default 1: with a halo, the ghost import runs on a side stream beside the elements that
   * hold no ghost node; 0 = in-stream before a single launch), "ctas_per_sm" and "scatter_mode" (tuning /
   * profiling only). */
  int b200s_set_option(b200s_ctx* ctx, const char* key, double value);

  /* ---- set-up (once per discretization) ------------------------------------------------------ */
  /* Device mesh = what Discretization::evaluate walks: the rank's COLUMN elements and COLUMN nodes
   * (4C_fem_discretization_evaluate.cpp:85, my_col_element_range).
   *   X    [host|dev] 3*n_col_nodes, node-major xyz       (Node::x(), calc_lib.hpp:161-190)
   *   conn [host|dev] n_col_eles*nodes_per_ele, column-node local ids in element node order
   * All elements of one context share cell type, formulation (displacement based, nonlinear TL)
   * and material; other element blocks use further contexts or stay on the CPU path.
   * n_col_eles == 0 (a processor without elements) and n_col_nodes == 0 (a processor that holds nothing) are valid, as are
   * n_row_nodes == 0 in b200s_build_graph / n_rows == 0 in b200s_set_graph (ghost elements only): the evaluate of such a
   * context has nothing to add, as the reference's column-element loop has nothing to do. */
  int b200s_set_mesh(b200s_ctx* ctx, int celltype, int64_t n_col_nodes, const double* X,
      int64_t n_col_eles, const int32_t* conn);

  /* Formulation of the element block = the SolidCalcVariant the reference's factory would create
   * (create_solid_calculation_interface, 4C_solid_ele_factory.cpp:53-55,252-305).  Default: nonlinear
   * total Lagrangean, no element technology, no prestress.  Returns B200S_ERR_UNSUPPORTED for a
   * combination without device kernel (the adapter keeps such blocks on the CPU element loop). */
  int b200s_set_element_properties(b200s_ctx* ctx, int kinem_type, int element_technology,
      int prestress_technology);

  /* Replaces the Mat::So3Material object(s) the elements hold (Mat::factory; So3Material::evaluate,
   * 4C_mat_so3_material.hpp:69-73).  per_ele_param: NULL or n_col_eles doubles overriding params[0]
   * element-wise (IsoNeoHooke MUE is an InputField looked up by element id,
   * 4C_mat_elast_isoneohooke.cpp:15-18,38). */
  int b200s_set_material(b200s_ctx* ctx, int kind, const double* params, int nparams,
      const double* per_ele_param);

  /* MAT_ElastHyper with NUMMAT summands (Mat::ElastHyper, src/mat/4C_mat_elasthyper.cpp:314-325;
   * elast_hyper_evaluate, 4C_mat_elasthyper_service.cpp:23-75).  kinds[n]: b200s_summand,
   * modes[n] (may be NULL = all 0), params: 4 doubles per summand (unused ones ignored).
   * B200S_ERR_UNSUPPORTED for an unknown summand, B200S_ERR_INVALID where the reference's
   * parameter constructor would throw. */
  int b200s_set_material_elasthyper(b200s_ctx* ctx, int n_summands, const int32_t* kinds,
      const int32_t* modes, const double* params);
  /* Element fibres (FIBER1 of the SOLID element line, Discret::Elements::read_fibers,
   * src/solid_ele/4C_solid_ele_fiber.cpp:45-73; Mat::Anisotropy::set_element_fibers): 3 doubles per
   * column element in the caller's order, normalised here like the reference does.  [host|dev]. */
  int b200s_set_element_fibers(b200s_ctx* ctx, const double* fiber1);
  /* Mat::So3Material::density() of the block (DENS): needed by the mass actions */
  int b200s_set_density(b200s_ctx* ctx, double density);
  /* EvaluationContext::time_step_size of the coming evaluates (4C_mat_visconeohooke.cpp:238-240): rate-dependent
   * materials only; without it their evaluate fails like the reference's FOUR_C_ASSERT. */
  int b200s_set_time_step_size(b200s_ctx* ctx, double dt);
  /* Row (owned) elements among the column elements: is_row[n_col_eles] in the caller's element
   * order, NULL = all.  Only struct_calc_energy distinguishes them (row elements only,
   * model_evaluator_structure.cpp:1714-1716). */
  int b200s_set_row_elements(b200s_ctx* ctx, const uint8_t* is_row);
  /* Discretization::set_state(name, vector on the dof row map) for the states the device path
   * reads: "displacement" (same as passing u_row to b200s_evaluate) and "acceleration"
   * (inertia force of the mass actions, get_acceleration_vector, 4C_solid_ele_evaluate.cpp:29-34). */
  int b200s_set_state(b200s_ctx* ctx, const char* name, const double* v_row);

  /* The reference's own sparsity graph after SparseMatrix::complete() (4C_linalg_sparsematrix.cpp:
   * 700-721): local CSR of the Epetra_CrsMatrix as returned by ExtractCrsDataPointers.
   *   rowptr      [host|dev] n_rows+1 (64-bit: > 2^31 nnz per rank is supported here)
   *   colind      [host|dev] nnz, matrix-local column ids, ascending inside each row
   *   dof_row_lid [host|dev] 3*n_col_nodes: rowmap.lid(gid) of each column-node dof, -1 if the
   *               dof is not owned by this rank (lmrowowner != myrank, :386)
   *   dof_col_lid [host|dev] 3*n_col_nodes: colmap.lid(gid) of each column-node dof (:374)
   * Builds the element -> CSR slot scatter map on the device with the same lower_bound + stride-3
   * contiguity rule as SparseMatrix::assemble (:404-466).  Returns B200S_ERR_UNSUPPORTED when a
   * node's three dofs are not contiguous rows/columns (the reference's slow path) or an element
   * column is missing from an owned row. */
  int b200s_set_graph(b200s_ctx* ctx, int64_t n_rows, const int64_t* rowptr, const int32_t* colind,
      const int32_t* dof_row_lid, const int32_t* dof_col_lid);

  /* Several element blocks, one system matrix.  Discretization::evaluate assembles ANY mix of cell types,
   * formulations and MAT ids into the same SparseMatrix (4C_fem_discretization_evaluate.cpp:85-117); a
   * context evaluates ONE block (cell type, SolidElementProperties, material).  b200s_attach_block makes
   * `ctx` a further block of `primary`'s matrix: after its own b200s_set_mesh (the SAME column nodes, in the
   * same order, as the primary; its own column elements), b200s_set_element_properties and material calls,
   * attach builds the block's element -> CSR slot map against the primary's graph (which must hold every
   * coupling of the block's elements: pass the matrix graph of the whole discretization to the primary's
   * b200s_set_graph) and from then on the block shares the primary's CSR values, residual, displacement
   * vectors, stream and communicator.  Per Newton iteration: b200s_evaluate on the primary first (state import,
   * zero, its elements), then on every attached block (its elements are ADDED; u_row is ignored / may be NULL);
   * K_values / f_row given to the LAST call receive the sum.  Destroy attached blocks before their primary. */
  int b200s_attach_block(b200s_ctx* ctx, b200s_ctx* primary);

  /* Ghost-displacement halo = the Epetra_Import of Discretization::set_state
   * (4C_fem_discretization.cpp:653-661), rank-local view (Epetra_Import::ExportLIDs / RemoteLIDs /
   * ExportPIDs).  For each neighbour n in [0,n_nbr): peer rank peer[n],
   *   send_row_lids[send_offsets[n]..send_offsets[n+1]) : ROW lid of dof 0 of each owned node whose
   *       three displacements go to that peer, in the order the peer expects them
   *   recv_col_lids[recv_offsets[n]..recv_offsets[n+1]) : COLUMN-node lids filled from that peer
   * [host] arrays.  The exchange itself is ncclSend/ncclRecv in one group on the context stream
   * between a pack and an unpack kernel; the communicator comes from b200s_comm_init_nccl. */
  int b200s_set_halo(b200s_ctx* ctx, int n_nbr, const int32_t* peer, const int64_t* send_offsets,
      const int32_t* send_row_lids, const int64_t* recv_offsets, const int32_t* recv_col_lids);
  /* Row-element mode = FE_MATRIX assembly (SparseMatrix::fe_assemble collects the rows a rank does not own, complete()
   * exports them to their owners in Add mode: 4C_linalg_sparsematrix.cpp:598-630,705-710).  The caller hands
   * b200s_set_mesh its ROW elements only and b200s_set_graph an EXTENDED graph: the rank's own dof rows (full pattern)
   * followed by "export rows" for the ghost nodes its row elements touch; before b200s_set_graph the options
   * "owned_rows" / "owned_nnz" (b200s_set_option) give the split.  u_row, K_values, f_row of b200s_evaluate keep their
   * meaning (own rows only).  After the element kernels the export rows travel:
   *   send_val / send_f [2 n_peers] : {start, count} of the slice of the extended values / residual for peer k (the
   *                                   export rows of one owner are contiguous), sent in place with ncclSend
   *   recv_pos[recv_offsets[k]..recv_offsets[k+1])  : for what peer k sends, in its order, the position in the rank's
   *                                   own CSR values to add it to; recv_fpos / recv_f_offsets likewise for the residual
   * = the plan of the Epetra_Export of the FE matrix ([host] arrays).  No peers: plain evaluation of the given elements.
   * The pipelined host copy, mass actions, Neumann loads and the device linear algebra are not offered in this mode. */
  int b200s_set_row_export(b200s_ctx* ctx, int n_peers, const int32_t* peers, const int64_t* send_val, const int64_t* send_f,
      const int64_t* recv_offsets, const int64_t* recv_pos, const int64_t* recv_f_offsets, const int32_t* recv_fpos);
  /* Host-mediated flavour of the same Add (tests without a communicator): adds export rows handed in by the caller --
   * values / f [host|dev] in the order of peer peer_index, either may be NULL -- at the positions of the plan. */
  int b200s_add_exported(b200s_ctx* ctx, int peer_index, const double* values, const double* f);
  /* NCCL bootstrap without MPI: rank 0 creates the 128-byte ncclUniqueId, the caller distributes
   * it (the harness uses torch.distributed; 4C would MPI_Bcast it), every rank initialises.
   * libnccl.so.2 is resolved at run time (dlopen), the library has no link-time NCCL dependency. */
  int b200s_nccl_unique_id(b200s_ctx* ctx, char* out128);
  int b200s_comm_init_nccl(b200s_ctx* ctx, const char* id128);

  /* ---- the hot path ---------------------------------------------------------------------------- */
  /* One Newton-iteration evaluate = Solid::ModelEvaluator::Structure::apply_force_stiff_internal
   * (model_evaluator_structure.cpp:442-472) restricted to the device element block:
   *   set_state("displacement") incl. ghost import      (4C_fem_discretization.cpp:612-665)
   *   stiff.zero(); fint.put_scalar(0)                  (:167-190; 4C_linalg_sparsematrix.cpp:268-291)
   *   for every column element: Solid<3>::evaluate(struct_calc_nlnstiff|internalforce) and
   *   AssembleStrategy::assemble_matrix1 / assemble_vector1 on owned rows
   *                                                     (4C_fem_discretization_evaluate.cpp:85-117)
   *   u_row    [host|dev] 3*n_row_nodes displacement on the dof ROW map
   *   K_values [host|dev] nnz doubles, OVERWRITTEN (zero + assemble); may be NULL for INTERNALFORCE
   *            or to leave K in the context's device buffer (b200s_device_values)
   *   f_row    [host|dev] n_rows doubles, OVERWRITTEN; may be NULL (stays on device)
   *   err      [host] optional
   * Returns B200S_OK, B200S_ERR_NEG_JACOBIAN (adapter rethrows the reference's message) ... */
  int b200s_evaluate(b200s_ctx* ctx, int action, const double* u_row, double* K_values,
      double* f_row, b200s_error_info* err);

  /* Results of the mass actions: second system matrix (same graph as K; mass_ in
   * structure_new) and the inertia force vector elevec2 assembled on the row map.  [host|dev] out. */
  int b200s_get_mass(b200s_ctx* ctx, double* M_values, double* f_inertia);
  double* b200s_device_mass_values(b200s_ctx* ctx); /* nnz, NULL before the first mass action */
  /* Copy of the CSR values of a system matrix of the context (b200s_matrix) as they are now, e.g.
   * after b200s_apply_dirichlet: what ExtractCrsDataPointers would show.  [host|dev] out, nnz. */
  int b200s_get_values(b200s_ctx* ctx, int which, double* values_out);
  double* b200s_device_f_inertia(b200s_ctx* ctx);   /* n_rows */
  /* Rows [row0,row1) of a resident system matrix (which: b200s_matrix) and of the residual: the values
   * Epetra_CrsMatrix::ExtractMyRowView (4C_linalg_sparsematrix.cpp:386-447 writes through it) would show
   * for those local rows, without moving the whole matrix -- the at-scale parity tests and bench.py's
   * `parity` record sample rows with it (BASELINE configs 3 and 5 hold > 2^31 non-zeros per rank).
   *   values_out [host|dev] rowptr[row1]-rowptr[row0] doubles, colind_out likewise (int32), rowptr_out
   *   row1-row0+1 offsets relative to the first returned entry; each may be NULL. */
  int b200s_get_value_rows(b200s_ctx* ctx, int which, int64_t row0, int64_t row1, double* values_out,
      int32_t* colind_out, int64_t* rowptr_out);
  int b200s_get_f_rows(b200s_ctx* ctx, int64_t row0, int64_t row1, double* f_out);

  /* struct_calc_stress (SolidEleCalc::calculate_stress, 4C_solid_ele_calc.cpp:416-476) for the
   * current "displacement" state: Gauss-point stresses / strains of every column element,
   * [n_col_eles][num_gp][6] doubles in the caller's element order, rows [xx,yy,zz,xy,yz,xz]
   * (what the reference serialises into stress_data / strain_data per element).  Either output
   * may be NULL.  [host|dev].  EAS blocks: the enhanced strain / stress, Cauchy and Euler-Almansi through
   * the enhanced deformation gradient F = R U_enh (4C_solid_ele_calc_lib.hpp:1566-1595). */
  int b200s_calc_stress(b200s_ctx* ctx, int stress_type, int strain_type, double* stress_out,
      double* strain_out);
  /* struct_calc_energy (calculate_internal_energy, 4C_solid_ele_calc.cpp:365-414): this rank's sum
   * of psi * detJ * w over its ROW elements (the caller does the sum_all,
   * model_evaluator_structure.cpp:1718-1725).  [host] out. */
  int b200s_calc_energy(b200s_ctx* ctx, double* energy_out);

  /* ---- MULF prestress (PRESTRESS_TECH mulf) --------------------------------------------------------- */
  /* struct_update_prestress (Solid<3>::evaluate -> SolidEleCalc::update_prestress,
   * 4C_solid_ele_calc.cpp:478-519; MulfFormulation::update_prestress, 4C_solid_ele_calc_mulf.hpp:
   * 124-143) for the current "displacement" state: every Gauss point stores the current deformation
   * gradient and pulls its inverse Jacobian forward.  The caller resets the displacements afterwards
   * (Solid::IMPLICIT::PreStress::post_update).  B200S_ERR_UNSUPPORTED (reference: FOUR_C_THROW) when
   * the block was not set up with B200S_PRESTRESS_MULF. */
  int b200s_update_prestress(b200s_ctx* ctx);
  /* Gauss-point history of the block = what MulfFormulation::pack / unpack write for restart
   * (4C_solid_ele_calc_mulf.hpp:112-122): per column element (caller's order) and Gauss point 18
   * doubles {inverse_jacobian[9], deformation_gradient[9]}, column-major 3x3 each.
   * b200s_gp_history_size: number of doubles (0 without history).  [host|dev]. */
  int64_t b200s_gp_history_size(const b200s_ctx* ctx);
  int b200s_get_gp_history(b200s_ctx* ctx, double* hist_out);
  int b200s_set_gp_history(b200s_ctx* ctx, const double* hist_in);

  /* ---- EAS (TECH eas_mild / eas_full): condensed element variables ------------------------------------ */
  /* struct_calc_recover (Solid::ModelEvaluator::Structure::run_recover, model_evaluator_structure.cpp:
   * 1482-1509; SolidEleCalc::recover default step, 4C_solid_ele_calc.cpp:284-314): after every Newton
   * update the caller passes the Newton direction ("residual displacement", dof row map, [host|dev]) and
   * the step length; every element updates alpha += step * (-Kaa^-1 (s + Kda^T du)) from the matrices its
   * last evaluate stored.  A no-op (B200S_OK) for formulations without condensed variables.  The
   * line-search correction (correct_condensed_variables_for_linesearch) is not on the device path. */
  int b200s_recover(b200s_ctx* ctx, const double* du_row, double step_length);
  /* Non-default step of struct_calc_recover (correct_condensed_variables_for_linesearch,
   * 4C_solid_ele_calc_eas.hpp:228-243): alpha += (new_step_length - old_step_length) * alpha_inc. */
  int b200s_recover_linesearch(b200s_ctx* ctx, double new_step_length);
  /* struct_calc_update_istep / struct_calc_reset_istep (SolidEleCalc::update / reset_to_last_converged,
   * 4C_solid_ele_calc.cpp:317-363,570-575): snapshot of the block's history at a converged step and the
   * way back after a failed one (before the first update_step: back to the fresh, default-constructed
   * history).  No-ops for formulations without history. */
  int b200s_update_step(b200s_ctx* ctx);
  int b200s_reset_step(b200s_ctx* ctx);
  /* Element history of the block = what EASFormulation::pack / unpack write for restart
   * (4C_solid_ele_calc_eas.hpp:266-288, without old_step_length): per column element (caller's order)
   * alpha_inc[n], alpha[n], s[n], invKaa[n*n], Kda[24*n], matrices column-major, n = 9 or 21. */
  int64_t b200s_element_history_size(const b200s_ctx* ctx);
  int b200s_get_element_history(b200s_ctx* ctx, double* hist_out);
  int b200s_set_element_history(b200s_ctx* ctx, const double* hist_in);
  /* EASHistoryData::old_step_length (4C_solid_ele_calc_eas.hpp:36,270,281): the reference packs it per
   * element, but every element holds the same value (the step length of the last recover), so the block
   * keeps one scalar; it is part of the update_step / reset_step snapshot and of the restart data. */
  double b200s_eas_old_step_length(const b200s_ctx* ctx);
  int b200s_set_eas_old_step_length(b200s_ctx* ctx, double step_length);

  /* ---- external loads: Discretization::evaluate_neumann restricted to the element block ------------- */
  /* (4C_fem_discretization_evaluate.cpp:264-297: every Line/Surface/VolumeNeumann condition, every
   * element of its geometry incl. ghosted ones, owner-computes assemble.)
   * One DESIGN SURF NEUMANN CONDITIONS entry on faces of the block's elements:
   *   face_nodes [host] n_faces * nodes_per_face column-node lids in the node order of the parent's
   *              surface (quad4 / quad9 / tri6 for hex8 / hex27 / tet10; the order the reference's
   *              boundary-element factory produces decides the sign of the orthopressure normal)
   *   onoff, val [host] 3 entries (ONOFF, VAL; the orthopressure types use entry 0 only and fail like
   *              the reference -- "orthopressure on 1st dof only!" -- otherwise)
   * Returns the condition id (>= 0) or a negative b200s_status. */
  int b200s_add_surface_neumann(b200s_ctx* ctx, int neumann_type, int64_t n_faces,
      const int32_t* face_nodes, const int32_t* onoff, const double* val);
  /* One DESIGN VOL NEUMANN CONDITIONS entry (dead body load, evaluate_neumann<celltype>,
   * src/solid_ele/4C_solid_ele_neumann_evaluator.cpp:43-113).  ele_ids [host] n_eles column-element
   * indices (set_mesh numbering) or NULL = every element of the block. */
  int b200s_add_volume_neumann(b200s_ctx* ctx, int64_t n_eles, const int32_t* ele_ids,
      const int32_t* onoff, const double* val);
  int b200s_clear_neumann(b200s_ctx* ctx);
  /* Functions of space and time stay with the host's function manager (FunctionOfSpaceTime::evaluate,
   * 4C_solid_ele_surface_evaluate.cpp:296-303,339-352): the caller sets, per condition,
   *   dof_fac [host] 3 factors (the value of the condition's time function per dof; NULL = 1) and
   *   gp_fac  [host|dev] n_gp*3 factors per Gauss point and dof (NULL = 1), evaluated at the positions
   *           b200s_neumann_gp_coords returns (material configuration; with B200S_NEUMANN_* flags the
   *           configuration an orthopressure condition acts on, taken from the resident displacement).
   * n_gp = b200s_neumann_gp_count = faces (or loaded elements) * Gauss points, face-major. */
  int64_t b200s_neumann_gp_count(const b200s_ctx* ctx, int cond);
  int b200s_neumann_gp_coords(b200s_ctx* ctx, int cond, int flags, double* xyz_out);
  int b200s_set_neumann_factors(b200s_ctx* ctx, int cond, const double* dof_fac, const double* gp_fac);
  /* Evaluate every condition: fext_row [host|dev] n_rows OVERWRITTEN with the external force
   * (may be NULL: stays in b200s_device_fext).  u_row [host|dev] displacement the deformation
   * dependent types act on ("displacement new" for orthopressure, the last converged "displacement"
   * for pseudo_orthopressure); NULL = the displacement resident from the last evaluate / set_state.
   * With B200S_NEUMANN_LOADLIN the linearisation of the orthopressure conditions is ADDED to the
   * stiffness values in the context, i.e. call it after the b200s_evaluate of the iteration
   * (apply_force_stiff_external after apply_force_stiff_internal) and before apply_dirichlet. */
  int b200s_evaluate_neumann(b200s_ctx* ctx, const double* u_row, int flags, double* fext_row);
  double* b200s_device_fext(b200s_ctx* ctx); /* n_rows, NULL before the first evaluate_neumann */

  /* ---- history materials: restart of the Gauss-point state store ------------------------------------- */
  /* Mat::PlasticDruckerPrager::pack / unpack (4C_mat_plasticdruckerprager.cpp:63-107): per column element
   * (set_mesh numbering) and Gauss point 14 doubles {strainpllast[6], strainbarpllast, strainplcurr[6],
   * strainbarplcurr}, tensor components 00,11,22,01,12,02.  Mat::ViscoNeoHooke::pack / unpack
   * (4C_mat_visconeohooke.cpp:80-158): 24 doubles {histstresslast[6], artstresslast[6], histstresscurr[6],
   * artstresscurr[6]}.  size = n_col_eles * num_gp * (14 | 24) doubles, 0 for a material without history.
   * [host|dev] */
  int64_t b200s_material_history_size(const b200s_ctx* ctx);
  int b200s_get_material_history(b200s_ctx* ctx, double* hist_out);
  int b200s_set_material_history(b200s_ctx* ctx, const double* hist_in);

  /* ---- the step after the path, on the device (K never returns to the host) ------------------------ */
  /* SparseMatrix::apply_dirichlet(dbctoggle, diagonalblock), in-place branch
   * (4C_linalg_sparsematrix.cpp:893-921): rows with toggle == 1.0 are zeroed and get 1.0 on the
   * diagonal if diagonalblock.  which: b200s_matrix.  dbctoggle_row [host|dev] n_rows. */
  int b200s_apply_dirichlet(b200s_ctx* ctx, int which, const double* dbctoggle_row, int diagonalblock);
  /* Core::LinAlg::apply_dirichlet_to_system(x, b, dbcval, dbctoggle)
   * (4C_linalg_utils_sparse_algebra_assemble.cpp:116-130); x_row or b_row may be NULL. */
  int b200s_apply_dirichlet_to_system(b200s_ctx* ctx, double* x_row, double* b_row,
      const double* dbcval_row, const double* dbctoggle_row);
  /* SparseMatrix::multiply(false, x, y): y = A x with x, y on the dof row map; the import of x to the
   * column map is the ghost exchange of b200s_set_halo.  [host|dev]. */
  int b200s_multiply(b200s_ctx* ctx, int which, const double* x_row, double* y_row);
  double b200s_last_multiply_ms(const b200s_ctx* ctx); /* device time of the last product */
  /* Hand-off demonstrator: block-Jacobi (3x3 node blocks) preconditioned CG on the stiffness matrix
   * in the context, x_row = initial guess in / solution out, stops at |r| <= rtol |r0|.
   * Not part of the reference path (4C hands K to Belos/MueLu); it shows the device CSR being
   * consumed in place.  Multi-rank: dot products by ncclAllReduce. */
  int b200s_solve_cg(b200s_ctx* ctx, const double* b_row, double* x_row, double rtol, int maxit,
      int* iters_out, double* relres_out);

  /* Asynchronous flavour used when the caller keeps everything on the device: enqueues on the
   * context stream, does not synchronise, u already in the context's device row vector. */
  int b200s_evaluate_async(b200s_ctx* ctx, int action);
  int b200s_synchronize(b200s_ctx* ctx, b200s_error_info* err);

  /* Device-resident buffers owned by the context (valid until destroy / next set_graph).
   * b200s_device_values: CSR values of the LAST stiffness evaluate (see option "double_buffer"). */
  double* b200s_device_values(b200s_ctx* ctx);  /* nnz  */
  double* b200s_device_f(b200s_ctx* ctx);       /* n_rows */
  double* b200s_device_u_row(b200s_ctx* ctx);   /* 3*n_row_nodes */
  double* b200s_device_u_col(b200s_ctx* ctx);   /* 3*n_col_nodes */
  void* b200s_stream(b200s_ctx* ctx);           /* cudaStream_t */
  int64_t b200s_nnz(const b200s_ctx* ctx);
  int64_t b200s_n_rows(const b200s_ctx* ctx);

  /* Timers (Teuchos::TimeMonitor stand-in, 4C_fem_discretization_evaluate.cpp:74): milliseconds of
   * the last evaluate measured with CUDA events on the context stream.
   * out[0]=halo, [1]=zero, [2]=element kernel(s), [3]=total device, [4]=h2d, [5]=d2h */
  int b200s_get_timers(b200s_ctx* ctx, double* out6);
  /* Number of kernels this library launched since create (bench's gpu_launches). */
  int64_t b200s_launch_count(const b200s_ctx* ctx);
  /* Accumulated device time (CUDA events on the context stream) and count of the element
   * evaluate+assemble kernel launches harvested so far (at b200s_synchronize); reset!=0 clears the
   * accumulators after reading.  bench.py's roofline figure is sum_ms/count over its timed region. */
  int b200s_kernel_time(b200s_ctx* ctx, double* sum_ms, int64_t* count, int reset);

  /* ---- standalone helpers (what 4C/Epetra provide in a real run) ------------------------------ */
  /* Builds AND installs the CSR graph a first assembly + FillComplete would produce for the mesh
   * in ctx (4C_linalg_sparsematrix.cpp:469-494,700-721) on the device (radix sort + unique of the
   * (row node, column node) pairs of all elements).  node_row_lid / node_col_lid [host|dev] are per
   * COLUMN NODE (row-node index or -1; matrix-local column-node index); dof lids are 3*lid+d.
   * rowptr_out (3*n_row_nodes+1) / colind_out (nnz) [host|dev] may be NULL; call once with
   * colind_out==NULL to learn *nnz_out, again to fetch colind.  A 4C run hands Epetra's graph to
   * b200s_set_graph instead; this helper is what the harness, tests and bench use. */
  int b200s_build_graph(b200s_ctx* ctx, int64_t n_row_nodes, const int32_t* node_row_lid,
      const int32_t* node_col_lid, int64_t* rowptr_out, int32_t* colind_out, int64_t* nnz_out);

  /* Roofline denominators measured on this device: FP64 FMA peak (TFLOP/s) of a register-resident
   * DFMA loop over the whole chip, and achieved clock. */
  int b200s_measure_fp64_peak(int device, double* tflops, double* sm_mhz);

  /* Diagnostics behind DESIGN.md's measurements (not on the hot path): which 0 DFMA TFLOP/s,
   * 1 DMMA m8n8k4 TFLOP/s, 2 DMMA+DFMA interleaved TFLOP/s, 3 cudaMemset GB/s over `size` bytes,
   * 4/5/6 red.global.add.f64 G atomics/s (warp-contiguous / 24-byte runs in an L2 window / random)
   * over a `size`-byte buffer. */
  int b200s_microbench(int device, int which, int64_t size, double* result);

  /* Pinned host memory for the adapter's staging of u / values / f (what the adapter copies into / out of
   * Epetra's own arrays, or registers them).  b200s_host_alloc_near places the pages on the NUMA node of the
   * given device (mmap + mbind + cudaHostRegister; *numa_node_out = that node, -1 when it fell back to a
   * plain cudaMallocHost): with one rank per GPU the K values of all ranks (GBs per Newton iteration) must
   * not land on one socket. */
  void* b200s_host_alloc(int64_t bytes);
  void* b200s_host_alloc_near(int device, int64_t bytes, int* numa_node_out);
  void b200s_host_free(void* p);
  /* Diagnostics of the last evaluate: out5[0] element-kernel launches, [1] chunks of K copied to the host
   * while later element segments were still being evaluated, [2] 1 if the ghost import overlapped the
   * elements without ghost nodes, [3] row phases of the installed graph, [4] elements holding a ghost node. */
  int b200s_last_evaluate_info(const b200s_ctx* ctx, int64_t* out5);
  /* Bytes of CSR values that crossed PCIe in the last b200s_evaluate with a host K_values, and whether the
   * block-upper transfer (option "symmetric_transfer") is in use for the installed graph (else: why not). */
  /* Diagnostics: the device element order, out[n_col_eles] = caller's column-element index of device element e. */
  int b200s_debug_element_order(b200s_ctx* ctx, int32_t* out);
  int64_t b200s_last_d2h_value_bytes(const b200s_ctx* ctx);
  const char* b200s_symmetric_transfer_status(const b200s_ctx* ctx);

  const char* b200s_version(void);

#ifdef __cplusplus
}
#endif
#endif
