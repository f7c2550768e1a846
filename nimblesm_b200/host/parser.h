// nimblesm_b200/host/parser.h — the input-deck surface: nimble::Parser / BlockProperties / IOFileName
// (src/nimble_parser.h:62-420, src/nimble_parser.cc:57-360) for the keys the explicit hex8 path reads.
// `key: value` lines, `#` comments; unknown keys throw std::invalid_argument exactly like the reference.
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace nimble_b200 {

// <base>.<label>.<extension>[.<num_ranks>.<zero padded rank>]  (src/nimble_parser.cc:57-88)
std::string
IOFileName(
    std::string const& serial_name,
    std::string const& extension,
    std::string const& label     = std::string(),
    int                my_rank   = 0,
    int                num_ranks = 0);

struct BlockProperties
{
  BlockProperties() : block_name_("none"), block_id_(-1) {}
  explicit BlockProperties(std::string props);  // "block_<id> <material key>"; id = suffix after the last '_'
  std::string block_name_;
  int         block_id_;
  std::string material_key_;
};

class Parser
{
 public:
  Parser();
  virtual ~Parser() = default;

  void
  Initialize();  // reads the file set with SetInputFilename
  void
  InitializeFromString(const std::string& deck_text);  // same grammar, text already in memory

  std::string
  GenesisFileName() const
  {
    return genesis_file_name_;
  }
  std::string
  ExodusFileName() const
  {
    return exodus_file_name_;
  }
  bool
  WriteTimingDataFile() const
  {
    return write_timing_data_file_;
  }
  std::string
  TimeIntegrationScheme() const;  // throws on anything but explicit / quasistatic
  double
  InitialTime() const
  {
    return initial_time_;
  }
  double
  FinalTime() const
  {
    return final_time_;
  }
  int
  NumLoadSteps() const
  {
    return num_load_steps_;
  }
  int
  OutputFrequency() const
  {
    return output_frequency_;
  }
  bool
  HasContact() const
  {
    return !contact_string_.empty();
  }
  std::string
  ContactString() const
  {
    return contact_string_;
  }
  // `contact visualization: visualize_contact_entities <on/off> visualize_bounding_boxes <on/off> file_name <name.e>`
  // (src/nimble_parser.cc:293-316, src/nimble_parser.h:195-206)
  bool
  ContactVisualization() const
  {
    return visualize_contact_entities_ || visualize_contact_bounding_boxes_;
  }
  std::string
  ContactVisualizationFileName() const
  {
    return contact_visualization_file_name_;
  }
  std::string
  GetModelMaterialParameters(int block_id) const;  // "none" for a block the deck does not list
  int
  GetBlockIdFromMaterial(const std::string& material_key) const;
  std::vector<std::string> const&
  GetBoundaryConditionStrings() const
  {
    return boundary_condition_strings_;
  }
  std::string
  GetOutputFieldString() const;  // throws when the deck has no "output fields"
  void
  SetRankID(int r)
  {
    my_rank_ = r;
  }
  int
  GetRankID() const
  {
    return my_rank_;
  }
  void
  SetNumRanks(int n)
  {
    num_ranks_ = n;
  }
  int
  GetNumRanks() const
  {
    return num_ranks_;
  }
  void
  SetInputFilename(const std::string& name)
  {
    file_name_ = name;
  }

 protected:
  virtual void
  ParseKeyValue(const std::string& key, const std::string& value);
  void
  ParseLine(std::string line);

  std::string                            file_name_{"none"};
  std::string                            genesis_file_name_{"none"}, exodus_file_name_{"none"};
  bool                                   use_two_level_mesh_decomposition_{false}, write_timing_data_file_{false};
  std::string                            time_integration_scheme_{"explicit"};
  double                                 nonlinear_solver_relative_tolerance_{1.0e-6};
  int                                    nonlinear_solver_max_iterations_{200};
  double                                 initial_time_{0.0}, final_time_{0.0};
  int                                    num_load_steps_{0}, output_frequency_{1};
  std::string                            contact_string_, contact_backend_string_, contact_visualization_string_;
  bool                                   visualize_contact_entities_ = false, visualize_contact_bounding_boxes_ = false;
  std::string                            contact_visualization_file_name_ = "none";
  std::map<std::string, std::string>     material_strings_;
  std::map<int, BlockProperties>         model_blocks_;
  std::vector<std::string>               boundary_condition_strings_;
  std::string                            output_field_string_;
  int                                    my_rank_{0}, num_ranks_{1};
};

}  // namespace nimble_b200
