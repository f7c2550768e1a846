// nimblesm_b200/host/genesis_mesh.h — Genesis (Exodus II on NetCDF-3) mesh reader with the accessor surface
// of nimble::GenesisMesh (src/nimble_genesis_mesh.h:56-360); the ex_* calls of GenesisMesh::ReadFile
// (src/nimble_genesis_mesh.cc:62-316) are replaced by direct NetCDF-3 reads (host/netcdf3.h).
#pragma once
#include <map>
#include <string>
#include <vector>

namespace nimble_b200 {

class GenesisMesh
{
 public:
  GenesisMesh() : file_name_("undefined"), dim_(-1) {}

  // Reads <file_name> (a serial .g or one Nemesis piece .g.<P>.<r>).  "none" leaves the mesh empty, like the
  // reference (src/nimble_genesis_mesh.cc:69).  Throws std::runtime_error on unreadable / malformed files.
  void
  ReadFile(std::string file_name);

  // In-memory construction (src/nimble_genesis_mesh.cc:429-478), used by the synthetic benchmark meshes.
  void
  Initialize(
      std::string const&                        file_name,
      std::vector<int> const&                   node_global_id,
      std::vector<double> const&                node_x,
      std::vector<double> const&                node_y,
      std::vector<double> const&                node_z,
      std::vector<int> const&                   elem_global_id,
      std::vector<int> const&                   block_ids,
      std::map<int, std::string> const&         block_names,
      std::map<int, std::vector<int>> const&    block_elem_global_ids,
      std::map<int, int> const&                 block_num_nodes_per_elem,
      std::map<int, std::vector<int>> const&    block_elem_connectivity,
      std::map<int, std::string> const&         node_set_names = {},
      std::map<int, std::vector<int>> const&    node_sets      = {});

  // Structured cube [0,1]^3 of n^3 unit-aspect hex8 in one block (id 1), node i + (n+1)(j + (n+1)k), Exodus
  // hex ordering; node sets 1 (all nodes) and 2 (x = 0 face): the synthetic mesh of SURVEY.md §8(d).
  static GenesisMesh
  StructuredCube(int n);

  bool
  IsValid() const
  {
    return file_name_ != "none";
  }
  std::string
  FileName() const
  {
    return file_name_;
  }
  unsigned int
  GetNumNodes() const
  {
    return (unsigned int)node_x_.size();
  }
  const int*
  GetNodeGlobalIds() const
  {
    return node_global_id_.data();
  }
  std::size_t
  GetNumNodeGlobalIds() const
  {
    return node_global_id_.size();
  }
  int
  GetMaxNodeGlobalId() const;
  unsigned int
  GetNumElements() const
  {
    return (unsigned int)elem_global_id_.size();
  }
  const int*
  GetElementGlobalIds() const
  {
    return elem_global_id_.data();
  }
  std::vector<int> const&
  GetElementGlobalIdsInBlock(int block_id) const
  {
    return block_elem_global_ids_.at(block_id);
  }
  unsigned int
  GetNumBlocks() const
  {
    return (unsigned int)block_ids_.size();
  }
  unsigned int
  GetNumGlobalBlocks() const
  {
    return (unsigned int)all_block_ids_.size();
  }
  bool
  HasBlock(std::string const& block_name) const;
  std::vector<int>
  GetBlockIds() const
  {
    return block_ids_;
  }
  std::vector<int>
  GetAllBlockIds() const
  {
    return all_block_ids_;
  }
  int
  GetNumElementsInBlock(int block_id) const;
  std::map<int, int>
  GetNumElementsInBlock() const;
  int
  GetNumNodesPerElement(int block_id) const
  {
    return block_num_nodes_per_elem_.at(block_id);
  }
  std::string
  GetElementType(int block_id) const;  // inferred from nodes per element (src/nimble_genesis_mesh.cc:480-505)
  std::string
  GetBlockName(int block_id) const
  {
    return all_block_names_.at(block_id);
  }
  int
  GetBlockId(std::string const& block_name) const;
  // ids of the named blocks that exist on this rank (src/nimble_genesis_mesh.cc:525-534)
  void
  BlockNamesToOnProcessorBlockIds(std::vector<std::string> const& block_names, std::vector<int>& block_ids) const;
  int
  GetDim() const
  {
    return dim_;
  }
  const double*
  GetCoordinatesX() const
  {
    return node_x_.data();
  }
  const double*
  GetCoordinatesY() const
  {
    return node_y_.data();
  }
  const double*
  GetCoordinatesZ() const
  {
    return node_z_.data();
  }
  const int*
  GetConnectivity(int block_id) const
  {
    return block_elem_connectivity_.at(block_id).data();
  }
  int
  GetNumNodeSets() const
  {
    return (int)node_set_ids_.size();
  }
  std::vector<int>
  GetNodeSetIds() const
  {
    return node_set_ids_;
  }
  std::map<int, std::string>
  GetNodeSetNames() const
  {
    return node_set_names_;
  }
  std::map<int, std::vector<int>>
  GetNodeSets() const
  {
    return node_sets_;
  }
  std::map<int, std::vector<double>>
  GetNodeSetDistributionFactors() const
  {
    return ns_distribution_factors_;
  }
  void
  Print(bool verbose = false, int my_rank = 0) const;

  // ---- in-driver domain decomposition (the B200 build's replacement for the offline SEACAS `decomp` step the
  //      reference needs before an MPI run, test/_wip/scaling_study/decomp.sh) ---------------------------------
  // Recursive coordinate bisection of the ELEMENTS by centroid: part of every element, blocks in GetBlockIds()
  // order, elements in block order.  Cuts fall on the longest extent; ties keep ascending element order, so every
  // rank computes the same partition from the same file.
  std::vector<int>
  RcbElementPartition(int n_parts) const;
  // Reduces this mesh to one part, exactly as a Nemesis piece presents it: the part's elements (block order kept),
  // every node they touch (ascending, shared nodes duplicated between parts and matched by global id,
  // src/nimble.mpi.reduction.cc:50-123), node sets restricted to those nodes, all block ids / names kept.
  void
  KeepPart(std::vector<int> const& part_of_element, int part);

 protected:
  std::string                        file_name_;
  int                                dim_;
  std::vector<int>                   node_global_id_;
  std::vector<double>                node_x_, node_y_, node_z_;
  std::vector<int>                   elem_global_id_;
  std::vector<int>                   block_ids_, all_block_ids_;
  std::map<int, std::string>         block_names_, all_block_names_;
  std::map<int, std::vector<int>>    block_elem_global_ids_;
  std::map<int, int>                 block_num_nodes_per_elem_;
  std::map<int, std::vector<int>>    block_elem_connectivity_;
  std::vector<int>                   node_set_ids_;
  std::map<int, std::string>         node_set_names_;
  std::map<int, std::vector<int>>    node_sets_;
  std::map<int, std::vector<double>> ns_distribution_factors_;
};

}  // namespace nimble_b200
