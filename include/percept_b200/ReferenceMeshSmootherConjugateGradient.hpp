// ReferenceMeshSmootherConjugateGradient.hpp — C++ host side of the B200-native smoother.
//
// Keeps the reference's smoother class API for this path so that a caller of
//   percept::ReferenceMeshSmootherConjugateGradientImpl<MeshType>
//   (src/percept/mesh/mod/smoother/ReferenceMeshSmootherConjugateGradient.hpp:74-174,
//    ReferenceMeshSmootherBase.hpp:28-105, MeshSmoother.hpp:35-80)
// can switch to percept_b200::ReferenceMeshSmootherConjugateGradientImpl<MeshType> with the same constructor
// arguments, the same public/virtual members and the same public state scalars.  All field arithmetic runs in the
// sm_100a kernels behind the C ABI (percept_smoother_c.h); this header only holds host control flow and converts
// non-zero statuses into std::runtime_error, the reference's error convention (Util.hpp:288-314).
//
// The mesh side is a minimal facade of the pieces of PerceptMesh / BlockStructuredGrid / STK the path touches:
// fields looked up by name, nodal BLAS-1, the string property map, boundary selectors evaluated once into a fixed
// mask.  Header-only; link with percept_b200/lib/libpms.so.
#ifndef PERCEPT_B200_REFERENCE_MESH_SMOOTHER_CG_HPP
#define PERCEPT_B200_REFERENCE_MESH_SMOOTHER_CG_HPP

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iomanip>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include "../percept_fixtures_c.h"
#include "../percept_smoother_c.h"

namespace percept_b200 {

typedef double Double;  // MeshType.hpp:41-46 under KOKKOS_ENABLE_CUDA

inline void check(pms_ctx* c, int rc) {
  if (rc != PMS_OK) throw std::runtime_error(std::string("percept_b200: ") + pms_last_error(c));
}

typedef std::array<unsigned, 4> StructuredCellIndex;  // i, j, k, block (structured/StructuredCellIndex.hpp)

// ---- structured side -------------------------------------------------------------------------------------------
struct StructuredBlock {  // structured/StructuredBlock.hpp:31-101 (sizes + boundary conditions + zone connectivity)
  std::array<unsigned, 3> node_size;
  struct BC {
    std::array<int, 3> beg, end;
  };
  struct ZoneConnectivity {
    int donor;
    std::array<int, 3> transform, owner_beg, owner_end, donor_beg, donor_end;
  };
  std::vector<BC> m_boundaryConditions;
  std::vector<ZoneConnectivity> m_zoneConnectivity;
  // host copy of the block coordinates (the reference's m_sgrid_coords), SoA planes x|y|z, i fastest; may be empty
  std::vector<double> m_sgrid_coords;
  size_t num_nodes() const { return (size_t)node_size[0] * node_size[1] * node_size[2]; }
  // StructuredBlock::dump_vtk (structured/StructuredBlock.cpp:297-445): one ASCII .vts piece, 4 decimals
  void dump_vtk(const std::string& file_prefix, unsigned iblock, size_t nblocks) const {
    std::ofstream out((file_prefix + "." + std::to_string(iblock) + ".vts").c_str());
    char buf[1024];
    std::snprintf(buf, sizeof buf,
                  "<VTKFile type=\"StructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
                  "  <StructuredGrid WholeExtent=\"0 %u 0 %u 0 %u\">\n"
                  "  <Piece Extent=\"0 %u 0 %u 0 %u\">\n"
                  "    <PointData>\n"
                  "    </PointData>\n",
                  node_size[0] - 1, node_size[1] - 1, node_size[2] - 1, node_size[0] - 1, node_size[1] - 1, node_size[2] - 1);
    out << buf << std::endl;
    out << "    <Points>\n"
           "       <DataArray type=\"Float64\" Name=\"Points\"  NumberOfComponents=\"3\" format=\"ascii\">\n";
    out.precision(4);
    out << std::fixed;
    const size_t n = num_nodes();
    for (size_t p = 0; p < n && m_sgrid_coords.size() == 3 * n; p++) {
      out << "         ";
      for (unsigned ic = 0; ic < 3; ic++) {
        double val = m_sgrid_coords[ic * n + p];
        if (std::fabs(val) < std::numeric_limits<double>::epsilon()) val = 0.0;
        out << std::setw(16) << val;
      }
      out << "\n";
    }
    out << "       </DataArray>\n"
           "    </Points>\n";
    out.unsetf(std::ios_base::floatfield);
    out.precision(6);
    const size_t ncell = (size_t)(node_size[0] - 1) * (node_size[1] - 1) * (node_size[2] - 1);
    out << "    <CellData Scalars=\" proc_id  block_id\">\n"
           "       <DataArray Name=\"proc_id\" type=\"Int32\" format=\"ascii\">\n         ";
    for (size_t e = 1; e <= ncell; e++) {
      out << " " << 0;
      if (e % 10 == 0) out << "\n         ";
    }
    out << "\n       </DataArray>\n"
           "       <DataArray Name=\"block_id\" type=\"Float32\" format=\"ascii\">\n         ";
    for (size_t e = 1; e <= ncell; e++) {
      out << " " << double(iblock) / double(nblocks);
      if (e % 10 == 0) out << "\n         ";
    }
    out << "\n       </DataArray>\n"
           "    </CellData>\n"
           "  </Piece>\n"
           "  </StructuredGrid>\n"
           "</VTKFile>\n"
        << std::endl;
  }
};

class PerceptMesh;

class BlockStructuredGrid {  // structured/BlockStructuredGrid.hpp:25-154
 public:
  std::vector<std::shared_ptr<StructuredBlock>> m_sblocks;
  // BlockStructuredGrid::fixture_1 / StructuredBlock::fixture_1 (StructuredBlock.cpp:101-168): one cube block, 6 BCs
  static std::shared_ptr<BlockStructuredGrid> fixture_1(std::array<unsigned, 3> sizes,
                                                        std::array<double, 3> widths = {{1, 1, 1}},
                                                        std::array<double, 3> offsets = {{0, 0, 0}}) {
    auto bsg = std::make_shared<BlockStructuredGrid>();
    auto b = std::make_shared<StructuredBlock>();
    b->node_size = sizes;
    const int n0 = (int)sizes[0], n1 = (int)sizes[1], n2 = (int)sizes[2];
    b->m_boundaryConditions = {{{{1, 1, 1}}, {{1, n1, n2}}},   {{{n0, 1, 1}}, {{n0, n1, n2}}}, {{{1, 1, 1}}, {{n0, 1, n2}}},
                               {{{1, n1, 1}}, {{n0, n1, n2}}}, {{{1, 1, 1}}, {{n0, n1, 1}}},   {{{1, 1, n2}}, {{n0, n1, n2}}}};
    b->m_sgrid_coords.resize(3 * b->num_nodes());
    pms_fixture_cube(sizes[0], sizes[1], sizes[2], widths.data(), offsets.data(), b->m_sgrid_coords.data());
    bsg->m_sblocks.push_back(b);
    return bsg;
  }
  // BlockStructuredGrid::dump_vtk / create_pvd (structured/BlockStructuredGrid.cpp:62-99, 185-200): .pvd + one .vts per block
  void dump_vtk(const std::string& file_prefix) const {
    std::ofstream out((file_prefix + ".pvd").c_str());
    out << "<VTKFile type=\"Collection\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n"
        << "  <Collection>\n";
    for (unsigned ib = 0; ib < m_sblocks.size(); ib++) {
      out << "      <DataSet part=\"" << ib << "\" file=\"" << file_prefix << "." << ib << ".vts\"/>" << std::endl;
      m_sblocks[ib]->dump_vtk(file_prefix, ib, m_sblocks.size());
    }
    out << "  </Collection>\n</VTKFile>\n";
  }
};

struct SGridSelector {  // MeshType.hpp:149-153
  virtual bool operator()(StructuredCellIndex& index) = 0;
  virtual ~SGridSelector() {}
};
struct SGridMeshGeometry {};  // MeshType.hpp:155-166 (stub in the reference too)

// ---- unstructured (STK) side: the facade takes plain arrays --------------------------------------------------------
struct STKSelector {  // stands for stk::mesh::Selector evaluated on node buckets: node -> selected?
  std::function<bool(size_t node)> pred;
  bool operator()(size_t node) const { return pred(node); }
};
// MeshGeometry (mesh/geometry/kernel/MeshGeometry.hpp:155-265): the reference classifies nodes against CAD entities
// (classify_node -> dof 0 vertex / 1 curve / 2 surface / 3 volume + evaluator) and evaluates normal_at /
// snap_points_to_geometry through an OpenNURBS / PGEOM kernel.  Here the geometry is a set of analytic surfaces
// evaluated on the GPU and the classification is handed over per node.
class MeshGeometry {
 public:
  struct Surface {
    int kind;
    double params[7];
  };
  int add_plane(const std::array<double, 3>& point, const std::array<double, 3>& normal) {
    return add(PMS_SURF_PLANE, {{point[0], point[1], point[2], normal[0], normal[1], normal[2], 0.0}});
  }
  int add_sphere(const std::array<double, 3>& centre, double radius) {
    return add(PMS_SURF_SPHERE, {{centre[0], centre[1], centre[2], radius, 0.0, 0.0, 0.0}});
  }
  int add_cylinder(const std::array<double, 3>& axis_point, const std::array<double, 3>& axis_dir, double radius) {
    return add(PMS_SURF_CYLINDER, {{axis_point[0], axis_point[1], axis_point[2], axis_dir[0], axis_dir[1], axis_dir[2], radius}});
  }
  // dof[n] as classify_node returns it, evaluator[n] = surface id of dof-2 nodes
  void set_classification(const std::vector<int8_t>& dof, const std::vector<int32_t>& evaluator) {
    m_dof = dof;
    m_evaluator = evaluator;
  }
  int classify_node(size_t node, size_t& curveOrSurfaceEvaluator) const {
    curveOrSurfaceEvaluator = node < m_evaluator.size() && m_evaluator[node] >= 0 ? (size_t)m_evaluator[node] : 0;
    return node < m_dof.size() ? (int)m_dof[node] : 3;
  }
  std::vector<Surface> m_surfaces;
  std::vector<int8_t> m_dof;
  std::vector<int32_t> m_evaluator;

 private:
  int add(int kind, const std::array<double, 7>& p) {
    Surface s;
    s.kind = kind;
    for (int i = 0; i < 7; i++) s.params[i] = p[i];
    m_surfaces.push_back(s);
    return (int)m_surfaces.size() - 1;
  }
};

struct StructuredGrid {
  typedef SGridSelector MTSelector;
  typedef SGridMeshGeometry MTMeshGeometry;
  typedef StructuredCellIndex MTNode;
  typedef StructuredCellIndex MTElement;
};
struct STKMesh {
  typedef STKSelector MTSelector;
  typedef MeshGeometry MTMeshGeometry;
  typedef size_t MTNode;
  typedef size_t MTElement;
};

// ---- PerceptMesh facade ----------------------------------------------------------------------------------------
class PerceptMesh {
 public:
  explicit PerceptMesh(int spatial_dim = 3, int device = 0) : m_spatial_dim(spatial_dim) {
    int rc = pms_create(&m_ctx, device);
    if (rc != PMS_OK) {
      std::string msg = m_ctx ? pms_last_error(m_ctx) : "pms_create failed";
      if (m_ctx) pms_destroy(m_ctx);
      m_ctx = nullptr;
      throw std::runtime_error("percept_b200: " + msg);
    }
  }
  ~PerceptMesh() {
    if (m_ctx) pms_destroy(m_ctx);
  }
  PerceptMesh(const PerceptMesh&) = delete;
  PerceptMesh& operator=(const PerceptMesh&) = delete;

  pms_ctx* ctx() const { return m_ctx; }
  int get_spatial_dim() const { return m_spatial_dim; }
  int get_rank() const { return m_rank; }
  // mesh_adapt --smooth_surfaces (PerceptMesh.hpp: get/set_smooth_surfaces): MS_SURFACE nodes slide on their surface
  bool get_smooth_surfaces() const { return m_smooth_surfaces; }
  void set_smooth_surfaces(bool v) {
    m_smooth_surfaces = v;
    check(m_ctx, pms_set_option(m_ctx, "smooth_surfaces", v ? 1.0 : 0.0));
  }

  // string property map (PerceptMesh.hpp:737-738)
  void setProperty(const std::string& k, const std::string& v) { m_props[k] = v; }
  std::string getProperty(const std::string& k) const {
    auto it = m_props.find(k);
    return it == m_props.end() ? std::string() : it->second;
  }

  // structured
  void set_block_structured_grid(std::shared_ptr<BlockStructuredGrid> bsg) {
    m_bsg = bsg;
    for (auto& b : bsg->m_sblocks) {
      int id = -1;
      check(m_ctx, pms_add_structured_block(m_ctx, b->node_size.data(), &id));
      for (auto& bc : b->m_boundaryConditions) check(m_ctx, pms_block_add_boundary_condition(m_ctx, id, bc.beg.data(), bc.end.data()));
    }
    for (size_t ib = 0; ib < bsg->m_sblocks.size(); ib++)
      for (auto& z : bsg->m_sblocks[ib]->m_zoneConnectivity)
        check(m_ctx, pms_block_add_zone_connectivity(m_ctx, (int)ib, z.donor, z.transform.data(), z.owner_beg.data(),
                                                     z.owner_end.data(), z.donor_beg.data(), z.donor_end.data()));
  }
  std::shared_ptr<BlockStructuredGrid> get_block_structured_grid() const { return m_bsg; }
  // unstructured (what PerceptMesh::open + STK hand to the smoother): connectivity in Exodus order
  void set_unstructured_mesh(int topology, size_t nnodes, size_t nelems, const int32_t* conn, const uint8_t* elem_owned = nullptr,
                             const uint8_t* node_owned = nullptr) {
    check(m_ctx, pms_set_unstructured(m_ctx, topology, nnodes, nelems, conn, elem_owned, node_owned));
    m_has_bulk = true;
  }
  bool get_bulk_data() const { return m_has_bulk; }

  // add_coordinate_state_fields (PerceptMesh.cpp:4490-4539) = commit: allocates the named fields on the device
  void add_coordinate_state_fields() {
    if (!m_committed) check(m_ctx, pms_commit(m_ctx));
    m_committed = true;
    if (m_bsg)  // the grid's own coordinates (fixture_1, a refiner's output, a reader) become the "coordinates" field
      for (size_t ib = 0; ib < m_bsg->m_sblocks.size(); ib++) {
        auto& b = *m_bsg->m_sblocks[ib];
        if (b.m_sgrid_coords.size() == 3 * b.num_nodes()) set_field("coordinates", b.m_sgrid_coords.data(), PMS_SOA, (int)ib);
      }
  }
  bool committed() const { return m_committed; }
  void set_fixed_mask(const uint8_t* mask) { check(m_ctx, pms_set_fixed_mask(m_ctx, mask)); }

  // field access by name
  void set_field(const std::string& name, const double* host, int layout = PMS_SOA, int block = -1) {
    check(m_ctx, pms_field_upload(m_ctx, name.c_str(), block, host, layout));
  }
  void get_field(const std::string& name, double* host, int layout = PMS_SOA, int block = -1) const {
    check(m_ctx, pms_field_download(m_ctx, name.c_str(), block, host, layout));
  }
  size_t num_nodes_local() const {
    size_t n = 0;
    pms_num_nodes_local(m_ctx, &n);
    return n;
  }

  // nodal BLAS-1 (PerceptMesh.hpp:459-536)
  void copy_field(const std::string& dst, const std::string& src) { check(m_ctx, pms_copy_field(m_ctx, dst.c_str(), src.c_str())); }
  void nodal_field_axpby(double alpha, const std::string& x, double beta, const std::string& y) {
    check(m_ctx, pms_nodal_field_axpby(m_ctx, alpha, x.c_str(), beta, y.c_str()));
  }
  void nodal_field_axpbypgz(double alpha, const std::string& x, double beta, const std::string& y, double gamma, const std::string& z) {
    check(m_ctx, pms_nodal_field_axpbypgz(m_ctx, alpha, x.c_str(), beta, y.c_str(), gamma, z.c_str()));
  }
  Double nodal_field_dot(const std::string& x, const std::string& y) {
    double r = 0;
    check(m_ctx, pms_nodal_field_dot(m_ctx, x.c_str(), y.c_str(), &r));
    return r;
  }
  void nodal_field_set_value(const std::string& x, double v) { check(m_ctx, pms_nodal_field_set_value(m_ctx, x.c_str(), v)); }
  int get_number_nodes() {
    size_t n = 0;
    check(m_ctx, pms_get_number_nodes(m_ctx, &n));
    return (int)n;
  }

 private:
  pms_ctx* m_ctx = nullptr;
  int m_spatial_dim, m_rank = 0;
  bool m_committed = false, m_has_bulk = false, m_smooth_surfaces = false;
  std::shared_ptr<BlockStructuredGrid> m_bsg;
  std::map<std::string, std::string> m_props;
};

// SGridBoundarySelector (ReferenceMeshSmootherConjugateGradient.hpp:31-69): node in any boundary-condition range
struct SGridBoundarySelector : public SGridSelector {
  PerceptMesh* m_eMesh;
  std::shared_ptr<BlockStructuredGrid> m_bsg;
  explicit SGridBoundarySelector(PerceptMesh* eMesh) : m_eMesh(eMesh), m_bsg(eMesh->get_block_structured_grid()) {}
  bool operator()(StructuredCellIndex& index) override {
    auto& sb = *m_bsg->m_sblocks[index[3]];
    for (auto& bc : sb.m_boundaryConditions) {
      bool in = true;
      for (int d = 0; d < 3; d++) {
        unsigned beg = (unsigned)std::min(bc.beg[d], bc.end[d]) - 1, end = (unsigned)std::max(bc.beg[d], bc.end[d]) - 1;
        in = in && beg <= index[d] && index[d] <= end;
      }
      if (in) return true;
    }
    return false;
  }
};

enum { MS_VERTEX, MS_CURVE, MS_SURFACE, MS_VOLUME, MS_ON_BOUNDARY, MS_NOT_ON_BOUNDARY };  // MeshType.hpp:29-38

// ---- StructuredGridRefiner (structured/StructuredGridRefiner.hpp:19-52, .cpp:40-170) -----------------------------
// 2x refinement of every block: do_refine() fills m_output (sizes 2n-1, boundary conditions and zone connectivity with
// the 1-based ranges mapped r -> 2r-1, refined coordinates).  The coordinates are prolonged on the GPU
// (pms_refine_structured_field); m_output carries a host copy like any other BlockStructuredGrid of this facade.
class StructuredGridRefiner {
 public:
  std::shared_ptr<BlockStructuredGrid> m_input, m_output;
  int m_debug;
  explicit StructuredGridRefiner(std::shared_ptr<BlockStructuredGrid> input, int debug = 0, int device = 0)
      : m_input(input), m_output(new BlockStructuredGrid()), m_debug(debug), m_device(device) {}

  unsigned do_refine() {
    struct Guard {  // two short-lived contexts: the coarse grid and its refinement
      pms_ctx* c = nullptr;
      ~Guard() {
        if (c) pms_destroy(c);
      }
    } ci, co;
    if (pms_create(&ci.c, m_device) != PMS_OK || pms_create(&co.c, m_device) != PMS_OK)
      throw std::runtime_error("percept_b200: StructuredGridRefiner: cannot create a device context");
    for (auto& b : m_input->m_sblocks) {
      int id = -1;
      check(ci.c, pms_add_structured_block(ci.c, b->node_size.data(), &id));
      for (auto& bc : b->m_boundaryConditions) check(ci.c, pms_block_add_boundary_condition(ci.c, id, bc.beg.data(), bc.end.data()));
    }
    for (size_t ib = 0; ib < m_input->m_sblocks.size(); ib++)
      for (auto& z : m_input->m_sblocks[ib]->m_zoneConnectivity)
        check(ci.c, pms_block_add_zone_connectivity(ci.c, (int)ib, z.donor, z.transform.data(), z.owner_beg.data(),
                                                    z.owner_end.data(), z.donor_beg.data(), z.donor_end.data()));
    check(ci.c, pms_commit(ci.c));
    for (size_t ib = 0; ib < m_input->m_sblocks.size(); ib++) {
      auto& b = *m_input->m_sblocks[ib];
      if (b.m_sgrid_coords.size() != 3 * b.num_nodes()) throw std::runtime_error("percept_b200: StructuredGridRefiner: input block has no coordinates");
      check(ci.c, pms_field_upload(ci.c, "coordinates", (int)ib, b.m_sgrid_coords.data(), PMS_SOA));
    }
    unsigned long long num_refined_cells = 0;
    check(co.c, pms_refine_structured_topology(ci.c, co.c, &num_refined_cells));
    check(co.c, pms_commit(co.c));
    check(co.c, pms_refine_structured_field(ci.c, "coordinates", co.c, "coordinates"));
    // the refined grid description (StructuredGridRefiner.cpp:55-154), m_index_base = 1
    m_output->m_sblocks.resize(0);
    auto map3 = [](const std::array<int, 3>& r) { return std::array<int, 3>{{2 * r[0] - 1, 2 * r[1] - 1, 2 * r[2] - 1}}; };
    for (size_t ib = 0; ib < m_input->m_sblocks.size(); ib++) {
      auto& b = *m_input->m_sblocks[ib];
      auto nb = std::make_shared<StructuredBlock>();
      for (int d = 0; d < 3; d++) nb->node_size[d] = 2 * b.node_size[d] - 1;
      for (auto& bc : b.m_boundaryConditions) nb->m_boundaryConditions.push_back({map3(bc.beg), map3(bc.end)});
      for (auto& z : b.m_zoneConnectivity)
        nb->m_zoneConnectivity.push_back({z.donor, z.transform, map3(z.owner_beg), map3(z.owner_end), map3(z.donor_beg), map3(z.donor_end)});
      nb->m_sgrid_coords.resize(3 * nb->num_nodes());
      check(co.c, pms_field_download(co.c, "coordinates", (int)ib, nb->m_sgrid_coords.data(), PMS_SOA));
      m_output->m_sblocks.push_back(nb);
    }
    return (unsigned)num_refined_cells;
  }

 private:
  int m_device;
};

// ---- smoother class hierarchy ------------------------------------------------------------------------------------
template <typename MeshType>
class MeshSmootherImpl {  // MeshSmoother.hpp:35-80
 protected:
  PerceptMesh* m_eMesh;
  int innerIter;
  double gradNorm;
  int parallelIterations;
  STKMesh::MTSelector* m_stk_boundarySelector;
  StructuredGrid::MTSelector* m_sgrid_boundarySelector;

 public:
  typename MeshType::MTMeshGeometry* m_meshGeometry;

  MeshSmootherImpl(PerceptMesh* eMesh, STKMesh::MTSelector* stk_select = 0, StructuredGrid::MTSelector* sgrid_select = 0,
                   typename MeshType::MTMeshGeometry* meshGeometry = 0, int innerIter_ = 100, double gradNorm_ = 1.e-8,
                   int parallelIterations_ = 20)
      : m_eMesh(eMesh), innerIter(innerIter_), gradNorm(gradNorm_), parallelIterations(parallelIterations_),
        m_stk_boundarySelector(stk_select), m_sgrid_boundarySelector(sgrid_select), m_meshGeometry(meshGeometry) {
  }
  virtual ~MeshSmootherImpl() {}
  StructuredGrid::MTSelector* get_sgrid_select() const { return m_sgrid_boundarySelector; }
  STKMesh::MTSelector* get_stkmesh_select() const { return m_stk_boundarySelector; }

  static size_t parallel_count_invalid_elements(PerceptMesh* eMesh) {  // MeshSmoother.cpp:192-234
    size_t n = 0;
    check(eMesh->ctx(), pms_count_invalid_elements(eMesh->ctx(), &n));
    return n;
  }
  void run() { this->run_algorithm(); }  // MeshSmoother.cpp:420-434
  virtual void run_algorithm() = 0;

  // get_fixed_flag (MeshSmoother.cpp:271-417) for both mesh types
  std::pair<bool, int> get_fixed_flag(StructuredCellIndex node) {
    if (m_sgrid_boundarySelector) {
      bool f = (*m_sgrid_boundarySelector)(node);
      return std::make_pair(f, f ? (int)MS_ON_BOUNDARY : (int)MS_NOT_ON_BOUNDARY);
    }
    return std::make_pair(false, (int)MS_VOLUME);
  }
  std::pair<bool, int> get_fixed_flag(size_t node) {
    if (m_stk_boundarySelector) {
      bool f = (*m_stk_boundarySelector)(node);
      return std::make_pair(f, f ? (int)MS_ON_BOUNDARY : (int)MS_NOT_ON_BOUNDARY);
    }
    return geometry_fixed_flag(m_meshGeometry, node);
  }

  /// project_delta_to_tangent_plane(node, delta, norm) (MeshSmoother.hpp:74, MeshSmoother.cpp:437-460): remove from
  /// delta its component along the surface normal at the node; norm (optional) receives the normal.  No geometry: no-op.
  /// The normal comes from the device evaluator (pms_get_surface_normals, cached until invalidate_normals()).
  void project_delta_to_tangent_plane(size_t node, double* delta, double* norm = 0) {
    if (!has_geometry(m_meshGeometry)) return;
    const size_t N = m_eMesh->num_nodes_local();
    if (m_normals.size() != 3 * N) {
      m_normals.assign(3 * N, 0.0);
      check(m_eMesh->ctx(), pms_get_surface_normals(m_eMesh->ctx(), m_normals.data()));
    }
    double dot = 0.0;
    for (int i = 0; i < m_eMesh->get_spatial_dim(); i++) dot += delta[i] * m_normals[i * N + node];
    for (int i = 0; i < m_eMesh->get_spatial_dim(); i++) {
      delta[i] -= dot * m_normals[i * N + node];
      if (norm) norm[i] = m_normals[i * N + node];
    }
  }
  void invalidate_normals() { m_normals.clear(); }
  /// snap_to(node, coordinate, reset) (MeshSmoother.hpp:79, MeshSmoother.cpp:517-563): if reset is true the node's
  /// coordinates are not modified and only the snapped position is returned in coordinate.
  void snap_to(size_t node, double* coordinate, bool reset = false) const {
    if (!has_geometry(m_meshGeometry)) return;
    check(m_eMesh->ctx(), pms_snap_point(m_eMesh->ctx(), node, coordinate, reset ? 1 : 0));
  }

 protected:
  std::vector<double> m_normals;
  static bool has_geometry(MeshGeometry* g) { return g != nullptr; }
  static bool has_geometry(SGridMeshGeometry*) { throw std::runtime_error("not impl"); }  // the reference's StructuredGrid specialisation
  // classification branch of get_fixed_flag (MeshSmoother.cpp:293-340)
  std::pair<bool, int> geometry_fixed_flag(MeshGeometry* geom, size_t node) {
    if (!geom) return std::make_pair(false, (int)MS_VOLUME);
    size_t evaluator = 0;
    const int dof = geom->classify_node(node, evaluator);
    if (dof == 0) return std::make_pair(true, (int)MS_VERTEX);
    if (dof == 1) return std::make_pair(true, (int)MS_CURVE);
    if (dof == 2) return std::make_pair(!m_eMesh->get_smooth_surfaces(), (int)MS_SURFACE);
    return std::make_pair(false, (int)MS_VOLUME);
  }
  std::pair<bool, int> geometry_fixed_flag(SGridMeshGeometry*, size_t) { return std::make_pair(false, (int)MS_VOLUME); }
  void upload_geometry(MeshGeometry* geom) {  // surfaces + classification -> device (fixed mask derived there)
    for (auto& s : geom->m_surfaces) check(m_eMesh->ctx(), pms_add_analytic_surface(m_eMesh->ctx(), s.kind, s.params, nullptr));
    const size_t N = m_eMesh->num_nodes_local();
    std::vector<int8_t> dof(N, 3);
    std::vector<int32_t> ev(N, -1);
    for (size_t n = 0; n < N; n++) {
      size_t e = 0;
      dof[n] = (int8_t)geom->classify_node(n, e);
      if (dof[n] == 2) ev[n] = (int32_t)e;
    }
    check(m_eMesh->ctx(), pms_set_node_classification(m_eMesh->ctx(), dof.data(), ev.data()));
  }
  void upload_geometry(SGridMeshGeometry*) {}

  // the selectors are evaluated once on the host into the device fixed mask
  void upload_fixed_mask() {
    if (!m_stk_boundarySelector && !m_sgrid_boundarySelector && m_meshGeometry && !m_eMesh->get_block_structured_grid()) {
      upload_geometry(m_meshGeometry);
      return;
    }
    const size_t N = m_eMesh->num_nodes_local();
    std::vector<uint8_t> mask(N, 0);
    auto bsg = m_eMesh->get_block_structured_grid();
    if (bsg) {
      size_t off = 0;
      for (unsigned ib = 0; ib < bsg->m_sblocks.size(); ib++) {
        auto& n = bsg->m_sblocks[ib]->node_size;
        for (unsigned k = 0; k < n[2]; k++)
          for (unsigned j = 0; j < n[1]; j++)
            for (unsigned i = 0; i < n[0]; i++) {
              StructuredCellIndex idx{{i, j, k, ib}};
              mask[off + i + (size_t)n[0] * (j + (size_t)n[1] * k)] = get_fixed_flag(idx).first ? 1 : 0;
            }
        off += bsg->m_sblocks[ib]->num_nodes();
      }
    } else {
      for (size_t n = 0; n < N; n++) mask[n] = get_fixed_flag(n).first ? 1 : 0;
    }
    m_eMesh->set_fixed_mask(mask.data());
  }
};

template <typename MeshType>
class ReferenceMeshSmootherBaseImpl : public MeshSmootherImpl<MeshType> {  // ReferenceMeshSmootherBase.hpp:28-105
 public:
  using Base = MeshSmootherImpl<MeshType>;
  ReferenceMeshSmootherBaseImpl(PerceptMesh* eMesh, STKMesh::MTSelector* stk_select = 0, StructuredGrid::MTSelector* sgrid_select = 0,
                                typename MeshType::MTMeshGeometry* meshGeometry = 0, int inner_iterations = 100,
                                double grad_norm = 1.e-8, int parallel_iterations = 20)
      : Base(eMesh, stk_select, sgrid_select, meshGeometry, inner_iterations, grad_norm, parallel_iterations) {
    if (eMesh->getProperty("ReferenceMeshSmootherBase_do_anim") != "") m_do_animation = std::stoi(eMesh->getProperty("ReferenceMeshSmootherBase_do_anim"));
    m_log_name = "smootherlog" + eMesh->getProperty("in_filename") + ".txt";  // ReferenceMeshSmootherBase.cpp:57
  }
  virtual ~ReferenceMeshSmootherBaseImpl() {}

 protected:
  virtual int get_num_stages() { return 2; }
  virtual double run_one_iteration() { throw std::runtime_error("not implemented"); }
  virtual bool check_convergence() { throw std::runtime_error("not implemented"); }
  std::string m_log_name;

 public:
  double m_scale = 0;
  Double m_dmax = 0, m_dmax_relative = 0;
  Double m_dnew = 0, m_dold = 0, m_d0 = 1, m_dmid = 0, m_dd = 0, m_alpha = 0, m_alpha_0 = 0, m_grad_norm = 0, m_grad_norm_scaled = 0;
  Double m_total_metric = 0;
  int m_stage = 0, m_iter = 0;
  size_t m_num_invalid = 0;
  Double m_global_metric = 1.7976931348623157e308;
  bool m_untangled = false;
  size_t m_num_nodes = 0;
  bool m_use_ref_mesh = true;
  int m_do_animation = 0;
};

template <typename MeshType>
class ReferenceMeshSmootherConjugateGradientImpl : public ReferenceMeshSmootherBaseImpl<MeshType> {
 public:
  using Base = ReferenceMeshSmootherBaseImpl<MeshType>;
  using Base::gradNorm;
  using Base::innerIter;
  using Base::m_alpha;
  using Base::m_alpha_0;
  using Base::m_d0;
  using Base::m_dmax;
  using Base::m_dmax_relative;
  using Base::m_dmid;
  using Base::m_dnew;
  using Base::m_dold;
  using Base::m_eMesh;
  using Base::m_global_metric;
  using Base::m_grad_norm_scaled;
  using Base::m_iter;
  using Base::m_num_invalid;
  using Base::m_num_nodes;
  using Base::m_stage;
  using Base::m_total_metric;
  using Base::m_untangled;
  using Base::m_use_ref_mesh;

  /// same constructor as the reference (ReferenceMeshSmootherConjugateGradient.hpp:121-128); pointers are borrowed
  ReferenceMeshSmootherConjugateGradientImpl(PerceptMesh* eMesh, STKMesh::MTSelector* stk_select = 0,
                                             StructuredGrid::MTSelector* sgrid_select = 0,
                                             typename MeshType::MTMeshGeometry* meshGeometry = 0, int inner_iterations = 100,
                                             double grad_norm = 1.e-8, int parallel_iterations = 20)
      : Base(eMesh, stk_select, sgrid_select, meshGeometry, inner_iterations, grad_norm, parallel_iterations),
        m_max_edge_length_factor(1.0) {
    if (!eMesh->committed()) eMesh->add_coordinate_state_fields();
    this->upload_fixed_mask();
  }
  virtual ~ReferenceMeshSmootherConjugateGradientImpl() {}

  double m_max_edge_length_factor;

  virtual void get_gradient() { check(ctx(), pms_get_gradient(ctx(), m_stage)); }  // Def.hpp:1481-1520

  virtual Double total_metric(Double alpha, double /*multiplicative_edge_scaling*/, bool& valid, size_t* num_invalid = 0) {  // Def.hpp:330-388
    sync_options();
    double m = 0;
    size_t n = 0;
    check(ctx(), pms_total_metric(ctx(), m_stage, alpha, &m, &n));
    valid = (n == 0);
    if (num_invalid) *num_invalid = n;
    return m;
  }
  virtual void update_node_positions(Double alpha) {  // Def.hpp:395-410
    double a = 0, b = 0;
    check(ctx(), pms_update_node_positions(ctx(), alpha, &a, &b));
    m_dmax = a;
    m_dmax_relative = b;
  }
  Double get_alpha_0() {  // Def.hpp:906-926
    double a = 0;
    check(ctx(), pms_get_alpha_0(ctx(), &a));
    return a;
  }
  void get_edge_lengths(PerceptMesh* /*eMesh*/) { check(ctx(), pms_get_edge_lengths(ctx())); }  // Def.hpp:759-771
  // get_surface_normals (Def.hpp:596-706): normals at non-fixed surface nodes, SoA (3, N), zero elsewhere
  std::vector<double> m_surface_normals;
  void get_surface_normals(PerceptMesh* eMesh) {
    m_surface_normals.assign(3 * eMesh->num_nodes_local(), 0.0);
    check(ctx(), pms_get_surface_normals(ctx(), m_surface_normals.data()));
  }
  void snap_nodes() {  // Def.hpp:273-316 (STKMesh); the StructuredGrid specialisation is empty in the reference too
    if (this->m_eMesh->get_block_structured_grid()) return;
    check(ctx(), pms_snap_nodes(ctx(), nullptr));
  }

  virtual bool check_convergence() {  // Def.hpp:989-1008
    pms_state st;
    to_state(st);
    return pms_check_convergence(&st, gradNorm) != 0;
  }

  /// line_search(restarted, mfac_mult) (ReferenceMeshSmootherConjugateGradient.hpp:155, Def.hpp:466-591): the accepted
  /// step length along cg_s; does not move the nodes.
  Double line_search(bool& restarted, double mfac_mult = 1.0) {
    if (dispatch_virtuals()) return line_search_host(restarted, mfac_mult);
    sync_options();
    pms_state st;
    to_state(st);
    int r = 0;
    double a = 0;
    check(ctx(), pms_line_search(ctx(), &st, mfac_mult, &r, &a));
    from_state(st);
    if (r) restarted = true;
    return a;
  }

  /// metric(element, valid) (ReferenceMeshSmootherConjugateGradient.hpp:169): the current stage's metric of one element.
  /// The device evaluates every element in one pass (there is no per-element host arithmetic); use
  /// pms_get_element_metrics directly for bulk access.
  virtual Double metric(size_t element, bool& valid) {
    sync_options();
    check(ctx(), pms_set_option(ctx(), "keep_element_metrics", 1.0));
    double m = 0;
    size_t n = 0;
    check(ctx(), pms_total_metric(ctx(), m_stage, 0.0, &m, &n));
    size_t ne = 0;
    check(ctx(), pms_num_elements_local(ctx(), &ne));
    std::vector<double> em(ne);
    std::vector<int32_t> ef(ne);
    check(ctx(), pms_get_element_metrics(ctx(), em.data(), ef.data()));
    check(ctx(), pms_set_option(ctx(), "keep_element_metrics", 0.0));
    if (element >= ne) throw std::runtime_error("metric: element out of range");
    valid = ef[element] == 0;
    return em[element];
  }
  /// nodal_edge_length_ave(node) (…ConjugateGradient.hpp:173, Def.hpp:39-65): mean over the node's elements of their mean
  /// edge length on the reference mesh = the node's cg_edge_length after get_edge_lengths().
  Double nodal_edge_length_ave(size_t node) {
    get_edge_lengths(m_eMesh);
    const size_t N = m_eMesh->num_nodes_local();
    std::vector<double> el(N);
    m_eMesh->get_field("cg_edge_length", el.data());
    if (node >= N) throw std::runtime_error("nodal_edge_length_ave: node out of range");
    return el[node];
  }

  // The exact class runs an iteration as ONE C-ABI call (line search, CG update and the fused passes stay on the device
  // side of the boundary).  A class DERIVED from this one may override get_gradient / total_metric /
  // update_node_positions / check_convergence, so for it the reference's control flow runs here on the host and calls
  // those virtuals, exactly as Def.hpp:1010-1093 / :466-591 / ReferenceMeshSmootherBase.cpp:240-334 do.
  virtual double run_one_iteration() {  // Def.hpp:1010-1093
    sync_options();
    if (dispatch_virtuals()) return run_one_iteration_host();
    pms_state st;
    to_state(st);
    double metric = 0;
    check(ctx(), pms_run_one_iteration(ctx(), &st, gradNorm, &metric));
    from_state(st);
    return metric;
  }

  virtual void run_algorithm() {  // ReferenceMeshSmootherBase.cpp:240-334 (writes smootherlog<in_filename>.txt)
    sync_options();
    if (dispatch_virtuals()) return run_algorithm_host();
    pms_state st;
    int rc = pms_run_algorithm(ctx(), innerIter, gradNorm, this->m_log_name.c_str(), &st);
    from_state(st);
    check(ctx(), rc);
  }

 protected:
  bool dispatch_virtuals() const { return typeid(*this) != typeid(ReferenceMeshSmootherConjugateGradientImpl<MeshType>); }

  Double line_search_host(bool& restarted, double mfac_mult) {  // Def.hpp:466-591 through the virtuals
    const Double alpha_fac = 10.0, tau = 0.5, c0 = 1.e-1, min_alpha_factor = 1.e-12;
    PerceptMesh* eMesh = m_eMesh;
    m_alpha_0 = get_alpha_0();
    Double alpha = alpha_fac * m_alpha_0;
    bool total_valid = false;
    size_t n_invalid = 0;
    const Double metric_0 = total_metric(0.0, 1.0, total_valid, &n_invalid);
    Double metric = 0.0;
    Double sDotGrad = eMesh->nodal_field_dot("cg_s", "cg_g");
    if (sDotGrad >= 0.0) {
      eMesh->copy_field("cg_s", "cg_r");
      m_alpha_0 = get_alpha_0();
      alpha = alpha_fac * m_alpha_0;
      sDotGrad = eMesh->nodal_field_dot("cg_s", "cg_g");
      restarted = true;
    }
    if (!(sDotGrad < 0.0)) throw std::runtime_error("bad sDotGrad");
    const Double armijo_offset_factor = c0 * sDotGrad;
    bool converged = false;
    total_valid = false;
    const int niter = 1000;
    for (int pass = 0; pass < 2 && !converged; pass++) {
      if (pass == 1) {  // second chance along the steepest-descent direction, Def.hpp:515-554
        restarted = true;
        eMesh->copy_field("cg_s", "cg_r");
        m_alpha_0 = get_alpha_0();
        alpha = alpha_fac * m_alpha_0;
        sDotGrad = eMesh->nodal_field_dot("cg_s", "cg_g");
        if (!(sDotGrad < 0.0)) throw std::runtime_error("bad sDotGrad 2nd time");
      }
      for (int liter = 0; !converged && liter < niter; ++liter) {
        metric = total_metric(alpha, 1.0, total_valid, &n_invalid);
        const Double mfac = alpha * armijo_offset_factor * mfac_mult;
        converged = (metric < metric_0 + mfac);
        if (m_untangled) {
          converged = converged && total_valid;
          if (pass == 1 && total_valid && m_dmax_relative < gradNorm) converged = true;
        }
        if (!converged) alpha *= tau;
        if (alpha < min_alpha_factor * m_alpha_0) break;
      }
    }
    if (!converged) throw std::runtime_error("can't reduce metric");
    const Double a1 = alpha / 2., a2 = alpha;
    const Double f0 = metric_0, f1 = total_metric(a1, 1.0, total_valid), f2 = total_metric(a2, 1.0, total_valid);
    const Double den = 2. * (a2 * (-f0 + f1) + a1 * (f0 - f2));
    const Double num = a2 * a2 * (f1 - f0) + a1 * a1 * (f0 - f2);
    if (std::fabs(den) > 1.e-10) {
      const Double alpha_quadratic = num / den;
      if (alpha_quadratic > 1.e-10 && alpha_quadratic < 2 * alpha) {
        const Double fm = total_metric(alpha_quadratic, 1.0, total_valid);
        if (fm < f2 && (m_stage == 0 || total_valid)) alpha = alpha_quadratic;
      }
    }
    return alpha;
  }

  double run_one_iteration_host() {  // Def.hpp:1010-1093 through the virtuals
    PerceptMesh* eMesh = m_eMesh;
    bool total_valid = false;
    if (m_num_nodes == 0) m_num_nodes = (size_t)eMesh->get_number_nodes();
    if (m_iter == 0) {
      get_edge_lengths(eMesh);
      eMesh->nodal_field_set_value("cg_r", 0.0);
      eMesh->nodal_field_set_value("cg_d", 0.0);
      eMesh->nodal_field_set_value("cg_s", 0.0);
    }
    get_gradient();
    eMesh->nodal_field_axpby(-1.0, "cg_g", 0.0, "cg_r");
    m_dold = eMesh->nodal_field_dot("cg_d", "cg_d");
    m_dmid = eMesh->nodal_field_dot("cg_r", "cg_d");
    m_dnew = eMesh->nodal_field_dot("cg_r", "cg_r");
    if (m_iter == 0) m_d0 = m_dnew;
    const Double metric_check = total_metric(0.0, 1.0, total_valid);
    m_total_metric = metric_check;
    if (check_convergence() || metric_check == 0.0) return total_metric(0.0, 1.0, total_valid);
    Double cg_beta = 0.0;
    if (m_dold == 0.0)
      cg_beta = 0.0;
    else if (m_iter > 0)
      cg_beta = (m_dnew - m_dmid) / m_dold;
    const size_t N = m_num_nodes;
    if (m_iter % N == 0 || cg_beta <= 0.0)
      eMesh->copy_field("cg_s", "cg_r");
    else
      eMesh->nodal_field_axpby(1.0, "cg_r", cg_beta, "cg_s");
    eMesh->copy_field("cg_d", "cg_r");
    bool restarted = false;
    const Double alpha = line_search(restarted);
    const Double snorm = eMesh->nodal_field_dot("cg_s", "cg_s");
    m_grad_norm_scaled = m_alpha_0 * std::sqrt(snorm) / Double(m_num_nodes);
    m_alpha = alpha;
    update_node_positions(alpha);
    if (eMesh->get_smooth_surfaces()) {
      snap_nodes();
      if (m_stage != 0) {
        bool total_valid_0 = true;
        total_metric(0.0, 1.0, total_valid_0);
        if (!total_valid_0) throw std::runtime_error("bad mesh after snap_node_positions...");
      }
    }
    return total_metric(0.0, 1.0, total_valid);
  }

  void run_algorithm_host() {  // ReferenceMeshSmootherBase.cpp:240-334 through the virtuals (no log file on this path)
    PerceptMesh* eMesh = m_eMesh;
    m_num_nodes = (size_t)eMesh->get_number_nodes();
    eMesh->copy_field("coordinates_lagged", "coordinates_NM1");
    eMesh->nodal_field_axpbypgz(1.0, "coordinates_N", 0.0, "coordinates_NM1", 0.0, "coordinates");
    m_num_invalid = Base::parallel_count_invalid_elements(eMesh);
    m_untangled = (m_num_invalid == 0);
    const int nstage = this->get_num_stages();
    for (int stage = 0; stage < nstage; stage++) {
      m_stage = stage;
      if (stage != 0 && Base::parallel_count_invalid_elements(eMesh) != 0)
        throw std::runtime_error("Invalid elements exist for start of stage 2, quiting");
      for (int iter = 0; iter < innerIter; ++iter) {
        m_iter = iter;
        m_num_invalid = Base::parallel_count_invalid_elements(eMesh);
        if (!m_untangled && m_num_invalid == 0) m_untangled = true;
        m_global_metric = this->run_one_iteration();
        const bool conv = this->check_convergence();
        if (!m_untangled && m_num_invalid == 0) m_untangled = true;
        if (conv && m_untangled) break;
      }
    }
    eMesh->copy_field("coordinates_lagged", "coordinates");
  }

 protected:
  pms_ctx* ctx() const { return m_eMesh->ctx(); }
  void sync_options() { check(ctx(), pms_set_option(ctx(), "use_ref_mesh", m_use_ref_mesh ? 1.0 : 0.0)); }
  void to_state(pms_state& st) {
    st = pms_state();
    st.dmax = m_dmax; st.dmax_relative = m_dmax_relative; st.dnew = m_dnew; st.dold = m_dold; st.d0 = m_d0; st.dmid = m_dmid;
    st.alpha = m_alpha; st.alpha_0 = m_alpha_0; st.grad_norm_scaled = m_grad_norm_scaled; st.total_metric = m_total_metric;
    st.global_metric = m_global_metric; st.stage = m_stage; st.iter = m_iter; st.untangled = m_untangled; st.num_invalid = m_num_invalid;
    if (m_num_nodes == 0) m_num_nodes = (size_t)m_eMesh->get_number_nodes();
    st.num_nodes = m_num_nodes;
  }
  void from_state(const pms_state& st) {
    m_dmax = st.dmax; m_dmax_relative = st.dmax_relative; m_dnew = st.dnew; m_dold = st.dold; m_d0 = st.d0; m_dmid = st.dmid;
    m_alpha = st.alpha; m_alpha_0 = st.alpha_0; m_grad_norm_scaled = st.grad_norm_scaled; m_total_metric = st.total_metric;
    m_global_metric = st.global_metric; m_stage = st.stage; m_iter = st.iter; m_untangled = st.untangled != 0; m_num_invalid = st.num_invalid;
    m_num_nodes = st.num_nodes;
  }
};

}  // namespace percept_b200
#endif
