// b200_solid_adapter.hpp -- host-side C++ layer above the C ABI (include/b200_solid.h).
//
// This is the piece a 4C maintainer links into the structure model evaluator: it owns one device
// context per (discretization, element block), translates 4C's maps / graph / vectors into the plain
// arrays the ABI takes and rethrows ABI status codes with the reference's own messages.  It is written
// against a small set of *views* (plain structs of pointers) instead of Epetra / 4C headers so that it
// compiles here without Trilinos; INTEGRATION.md shows the few lines that fill the views from
// Core::FE::Discretization, Core::LinAlg::SparseMatrix and Core::LinAlg::Vector<double>.
//
// Interface names follow the reference:
//   set_state / evaluate        Core::FE::Discretization (4C_fem_discretization.cpp:612-665,
//                                4C_fem_discretization_evaluate.cpp:67-117)
//   Action                      Core::Elements::ActionType struct_calc_nlnstiff / _internalforce
//   assemble semantics          Core::FE::AssembleStrategy::assemble_matrix1 / assemble_vector1
//                                (4C_fem_general_assemblestrategy.cpp:116-150): owned rows only,
//                                values ADDED into a zeroed matrix
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b200_solid.h"

namespace B200Solid
{
  enum class Action
  {
    struct_calc_internalforce = B200S_ACTION_INTERNALFORCE,
    struct_calc_nlnstiff = B200S_ACTION_NLNSTIFF,
    struct_calc_nlnstiffmass = B200S_ACTION_NLNSTIFFMASS,
    struct_calc_nlnstifflmass = B200S_ACTION_NLNSTIFFLMASS,
    struct_calc_internalinertiaforce = B200S_ACTION_INTERNALINERTIAFORCE
  };

  /// What 4C throws: Core::Exception (FOUR_C_THROW).  Stand-in so that unit tests can
  /// EXPECT_THROW on the same conditions as the reference (unittests/README.md).
  class Exception : public std::runtime_error
  {
   public:
    Exception(int code, const std::string& what) : std::runtime_error(what), code_(code) {}
    int code() const { return code_; }

   private:
    int code_;
  };

  /// Column-map view of one element block of a discretization (all elements share cell type,
  /// SolidElementProperties and material type) -- Discretization::my_col_element_range().
  struct ElementBlockView
  {
    int celltype = B200S_HEX8;        // Core::FE::CellType of the block
    std::int64_t n_col_nodes = 0;
    const double* X = nullptr;        // 3*n_col_nodes, Node::x() in node column-map order
    std::int64_t n_col_eles = 0;
    const std::int32_t* conn = nullptr;  // n_col_eles*nn, node column-map lids (Element::node_ids -> lid)
    const std::int32_t* dof_row_lid = nullptr;  // 3*n_col_nodes, dof_row_map()->lid(gid) or -1
    const std::int32_t* dof_col_lid = nullptr;  // 3*n_col_nodes, matrix col_map().lid(gid)
  };

  /// Epetra_CrsMatrix::ExtractCrsDataPointers of the completed stiffness matrix.  `values` changes
  /// every Newton iteration (SparseMatrix::zero() re-allocates, 4C_linalg_sparsematrix.cpp:268-291),
  /// the structure does not.
  struct CrsView
  {
    std::int64_t n_rows = 0;
    const int* index_offset = nullptr;  // n_rows+1 (Epetra: int)
    const int* indices = nullptr;       // nnz, local column ids, ascending per row
    double* values = nullptr;           // nnz
  };

  /// Epetra_Import of the dof column map (Discretization::set_state), per neighbour.
  struct ImportView
  {
    std::vector<int> peers;
    std::vector<std::vector<std::int32_t>> send_row_lids;  // row lid of dof 0 of each exported node
    std::vector<std::vector<std::int32_t>> recv_col_nodes; // column-node lids filled from the peer
  };

  // Export plan of an FE_MATRIX (Epetra_Export of the non-local rows, Add mode: 4C_linalg_sparsematrix.cpp:598-630):
  // what b200s_set_row_export takes, per peer.  The block is then set up with the rank's ROW elements and the extended
  // graph (own rows + export rows); set_option("owned_rows" / "owned_nnz") before set_graph.
  struct RowExportView
  {
    std::vector<int> peers;
    std::vector<std::int64_t> send_val_start, send_val_count;  // slice of the extended CSR values for the peer
    std::vector<std::int64_t> send_f_start, send_f_count;      // slice of the extended residual
    std::vector<std::vector<std::int64_t>> recv_pos;           // where the peer's export values are added (own CSR values)
    std::vector<std::vector<std::int32_t>> recv_fpos;          // ... and its residual entries
  };

  class DeviceElementBlock
  {
   public:
    DeviceElementBlock(int device, int rank, int nranks);
    ~DeviceElementBlock();
    DeviceElementBlock(const DeviceElementBlock&) = delete;
    DeviceElementBlock& operator=(const DeviceElementBlock&) = delete;

    /// true if (cell type, material) are on the device path; otherwise the caller keeps the CPU loop
    static bool supported(int celltype, int material_kind);
    /// The full capability matrix of the device path = which SolidCalcVariant / So3Material combinations of the
    /// reference's factory (4C_solid_ele_factory.cpp:252-305) have a kernel: cell type x KINEM x TECH x
    /// PRESTRESS_TECH x material kind.  Anything else keeps the CPU element loop.
    static bool supported(int celltype, int kinem, int tech, int prestress, int material_kind);
    /// This block assembles into `primary`'s system matrix (several cell types / MAT ids in one discretization,
    /// 4C_fem_discretization_evaluate.cpp:85-117): call after set_block / set_element_properties / set_material
    /// of this block and after the primary's set_graph with the graph of the WHOLE matrix.  Per Newton
    /// iteration evaluate the primary first, then the attached blocks; pass crs_values / f_row to the last one.
    void attach_to(DeviceElementBlock& primary);
    /// Pinned staging buffer on the NUMA node of this block's GPU (for Epetra value arrays the adapter copies
    /// into; b200s_host_alloc_near)
    static double* alloc_staging(int device, std::int64_t n_doubles);
    static void free_staging(double* p);

    void set_block(const ElementBlockView& block);
    /// Mat::So3Material of the block: kind + parameters (+ optional per-element field, see
    /// 4C_mat_elast_isoneohooke.cpp:15-18)
    void set_material(int kind, const std::vector<double>& params, const double* per_element = nullptr);
    /// MAT_ElastHyper with a summand list (b200s_summand ids, 4 input parameters each) and DENS
    void set_material_elasthyper(const std::vector<std::int32_t>& kinds, const std::vector<std::int32_t>& modes,
        const std::vector<double>& params4, double density = 0.0);
    /// Discret::Elements::SolidElementProperties of the block (KINEM / TECH / PRESTRESS_TECH)
    void set_element_properties(int kinem, int tech, int prestress);
    void set_row_elements(const std::uint8_t* is_row);
    void set_graph(const ElementBlockView& block, const CrsView& crs);
    void set_import(const ImportView& imp);
    void set_row_export(const RowExportView& exp);  // row-element mode: SparseMatrix::fe_assemble + complete()
    void init_communicator(const char* nccl_id128);  // id from rank 0's nccl_unique_id(), MPI_Bcast by 4C
    std::string nccl_unique_id();

    /// discret().set_state("displacement", u): u on the dof ROW map (host pointer)
    void set_state(const double* u_row) { u_row_ = u_row; }
    /// discret().evaluate(p{action}, stiff, nullptr, fint, ...): adds nothing on the host; overwrites
    /// crs.values (zero + assemble) and f_row with the device result.  Throws Exception with the
    /// reference's message text on det J <= 0.
    void evaluate(Action action, double* crs_values, double* f_row);
    /// further states by name, e.g. "acceleration" (mass actions)
    void set_state(const char* name, const double* v_row);
    /// results of the mass actions: mass_->values, inertia force
    void get_mass(double* mass_values, double* f_inertia);
    /// struct_calc_stress / struct_calc_energy / struct_update_prestress / struct_calc_recover
    void calc_stress(int stress_type, int strain_type, double* stress, double* strain);
    double calc_energy();
    void update_prestress();
    void recover(const double* du_row, double step_length);
    /// Discretization::evaluate_neumann on the block (4C_fem_discretization_evaluate.cpp:264-297):
    /// conditions of the block's faces / elements; functions of space and time are evaluated by the
    /// caller's function manager and handed in as factors (set_neumann_factors).
    int add_surface_neumann(b200s_neumann_type type, std::int64_t n_faces, const std::int32_t* face_nodes,
        const std::int32_t* onoff, const double* val);
    int add_volume_neumann(std::int64_t n_eles, const std::int32_t* ele_ids, const std::int32_t* onoff,
        const double* val);
    void set_neumann_factors(int cond, const double* dof_fac, const double* gp_fac);
    /// fext_row overwritten; loadlin adds the orthopressure load stiffness to the resident stiffness
    void evaluate_neumann(double* fext_row, bool loadlin, bool mulf_active = false, const double* u_row = nullptr);
    /// struct_calc_update_istep / struct_calc_reset_istep (element and material history) and restart of the
    /// material state store (Mat::PlasticDruckerPrager::pack / unpack)
    void update_step();
    void reset_step();
    std::int64_t material_history_size() const;
    void get_material_history(double* hist) ;
    void set_material_history(const double* hist);
    /// SparseMatrix::apply_dirichlet / multiply on the device-resident stiffness
    void apply_dirichlet(const double* dbctoggle_row, bool diagonalblock = true);
    void multiply(const double* x_row, double* y_row);

    std::int64_t nnz() const;
    std::int64_t n_rows() const;
    b200s_ctx* context() { return ctx_; }

   private:
    void check(int rc) const;
    b200s_ctx* ctx_ = nullptr;
    const double* u_row_ = nullptr;
    bool attached_ = false;
  };
}  // namespace B200Solid
