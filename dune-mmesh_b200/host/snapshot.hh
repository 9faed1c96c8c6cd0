// snapshot.hh -- what a host adapter extracts from the CGAL triangulation of a live Dune::MMesh (see include/mmesh_b200.h for
// the conventions), and the exception types of the host side.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace Dune {
namespace B200 {

// Dune::GridError analogue (mmesh.hh:1104-1107 and friends)
struct GridError : std::runtime_error { using std::runtime_error::runtime_error; };
struct InvalidStateException : std::runtime_error { using std::runtime_error::runtime_error; };

template <int dim>
using GlobalCoordinate = std::array<double, dim>;   // same layout as Dune::FieldVector<double, dim>

// What the adapter extracts from the CGAL triangulation (see include/mmesh_b200.h for the conventions).
template <int dim>
struct Snapshot {
  std::vector<GlobalCoordinate<dim>> x;             // vh->point(), leaf vertex index order
  std::vector<std::array<int32_t, dim + 1>> cells;  // face->vertex(k)->info().index, TDS order
  std::vector<uint8_t> perm;                        // ElementInfo::cgalIndex packed
  std::vector<std::array<int32_t, 4>> edges;        // (a,b,c,d), 2-D only
  std::vector<std::array<int32_t, dim>> facets;     // interface facets
  std::vector<uint8_t> vflags;                      // bit0 isInterface, bit1 boundaryFlag == 1 (host-side use only)
  std::vector<int32_t> interfaceVertices;           // interface leaf vertex index -> bulk vertex index (interface/indexsets.hh:238-277)
  std::vector<std::array<int32_t, 3>> neighbours;   // 2-D, optional: face->neighbor(k) cell index, -1 = infinite face
  std::vector<int32_t> insertionLevel;              // optional: VertexInfo::insertionLevel (mmesh.hh:132), default 0
  std::vector<int64_t> vertexId;                    // optional: VertexInfo::id (cgal/defaults.hh:41), persistent across TDS edits;
                                                    // needed by adapt(handle) to match cells of the old and the new snapshot
};

}  // namespace B200
}  // namespace Dune
