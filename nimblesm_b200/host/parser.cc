// nimblesm_b200/host/parser.cc — see parser.h.
#include "parser.h"

#include <algorithm>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

namespace nimble_b200 {

std::string
IOFileName(std::string const& serial_name, std::string const& extension, std::string const& label, int my_rank, int num_ranks)
{
  if (serial_name == "none") return serial_name;
  std::string base = serial_name;
  for (const char* suffix : {".g", ".e"}) {
    size_t pos = base.rfind(suffix);
    if (pos != std::string::npos) base = base.substr(0, pos);
  }
  if (!label.empty()) base += "." + label;
  base += "." + extension;
  if (num_ranks > 1) {
    const std::string total = std::to_string(num_ranks), mine = std::to_string(my_rank);
    base += "." + total + "." + std::string(total.size() - mine.size(), '0') + mine;
  }
  return base;
}

BlockProperties::BlockProperties(std::string props) : block_id_(-1)
{
  const size_t space = props.find(' ');
  block_name_        = props.substr(0, space);
  material_key_      = space == std::string::npos ? std::string() : props.substr(space + 1);
  // the material key is everything after the first blank, trimmed
  const size_t b = material_key_.find_first_not_of(" \t");
  material_key_  = b == std::string::npos ? std::string() : material_key_.substr(b);
  const size_t us = block_name_.rfind('_');
  std::istringstream(block_name_.substr(us == std::string::npos ? 0 : us + 1)) >> block_id_;
}

Parser::Parser() = default;

void
Parser::Initialize()
{
  std::ifstream fin(file_name_.c_str());
  if (!fin.good()) throw std::invalid_argument("\n**** Error in Parser::ReadFile(), unable to read file " + file_name_ + "\n");
  std::string line;
  while (std::getline(fin, line)) ParseLine(line);
}

void
Parser::InitializeFromString(const std::string& deck_text)
{
  std::istringstream in(deck_text);
  std::string        line;
  while (std::getline(in, line)) ParseLine(line);
}

void
Parser::ParseLine(std::string line)
{
  const size_t pound = line.find('#');
  if (pound != std::string::npos) line = line.substr(0, pound);
  auto trim = [](const std::string& s) {
    const size_t b = s.find_first_not_of(" \t\r");
    if (b == std::string::npos) return std::string();
    const size_t e = s.find_last_not_of(" \t\r");
    return s.substr(b, e - b + 1);
  };
  line = trim(line);
  if (line.empty()) return;
  const size_t colon = line.find(':');
  if (colon == std::string::npos) throw std::invalid_argument("\n**** Error in Parser::ReadFile(), unknown key " + line + "\n");
  ParseKeyValue(trim(line.substr(0, colon)), trim(line.substr(colon + 1)));
}

namespace {
bool
parse_switch(const std::string& key, const std::string& value)
{
  std::string up(value);
  std::transform(up.begin(), up.end(), up.begin(), [](unsigned char c) { return (char)std::toupper(c); });
  if (up == "TRUE" || up == "YES" || up == "ON") return true;
  if (up == "FALSE" || up == "NO" || up == "OFF") return false;
  throw std::invalid_argument("\n**** Error in Parser::ReadFile(), unexpected value for \"" + key + "\" " + value + "\n");
}
}  // namespace

void
Parser::ParseKeyValue(const std::string& key, const std::string& value)
{
  if (key == "genesis input file") {
    genesis_file_name_ = value;
  } else if (key == "exodus output file") {
    exodus_file_name_ = value;
  } else if (key == "use two level mesh decomposition") {
    use_two_level_mesh_decomposition_ = parse_switch(key, value);
  } else if (key == "write timing data file") {
    write_timing_data_file_ = parse_switch(key, value);
  } else if (key == "time integration scheme") {
    time_integration_scheme_ = value;
  } else if (key == "nonlinear solver relative tolerance") {
    nonlinear_solver_relative_tolerance_ = std::atof(value.c_str());
  } else if (key == "nonlinear solver maximum iterations") {
    nonlinear_solver_max_iterations_ = std::atoi(value.c_str());
  } else if (key == "initial time") {
    initial_time_ = std::atof(value.c_str());
  } else if (key == "final time") {
    final_time_ = std::atof(value.c_str());
  } else if (key == "number of load steps") {
    num_load_steps_ = std::atoi(value.c_str());
  } else if (key == "output frequency") {
    output_frequency_ = std::atoi(value.c_str());
  } else if (key == "contact") {
    contact_string_ = value;
  } else if (key == "contact backend") {
    contact_backend_string_ = value;
  } else if (key == "contact visualization") {
    contact_visualization_string_ = value;
    std::istringstream       ss(value);
    std::vector<std::string> vals;
    for (std::string val; ss >> val;) vals.push_back(val);
    if (vals.size() != 6 || vals[0] != "visualize_contact_entities" || (vals[1] != "on" && vals[1] != "off") ||
        vals[2] != "visualize_bounding_boxes" || (vals[3] != "on" && vals[3] != "off") || vals[4] != "file_name")
      throw std::invalid_argument(
          "\n**** Error in Parser::ReadFile(), unexpected value for \"contact visualization\"\n"
          "**** Allowable syntax is \"visualize_contact_entities <on/off> visualize_bounding_boxes <on/off> file_name <file_name.e>\"\n");
    visualize_contact_entities_       = vals[1] == "on";
    visualize_contact_bounding_boxes_ = vals[3] == "on";
    contact_visualization_file_name_  = vals[5];
  } else if (key == "material parameters") {
    const size_t space              = value.find(' ');
    material_strings_[value.substr(0, space)] = space == std::string::npos ? std::string() : value.substr(space + 1);
  } else if (key == "element block") {
    BlockProperties props(value);
    model_blocks_[props.block_id_] = props;
  } else if (key == "boundary condition") {
    boundary_condition_strings_.push_back(value);
  } else if (key == "output fields") {
    output_field_string_ = value;
  } else if (key == "contact dicing" || key == "contact splitting") {
    std::cout << " **** Parser::ReadFile(), skipping key " + key + "\n";
  } else {
    throw std::invalid_argument("\n**** Error in Parser::ReadFile(), unknown key " + key + "\n");
  }
}

std::string
Parser::TimeIntegrationScheme() const
{
  if (time_integration_scheme_ != "explicit" && time_integration_scheme_ != "quasistatic")
    throw std::invalid_argument("\n**** Error in Parser::TimeIntegrationScheme(), invalid integration scheme " +
                                time_integration_scheme_ + ".\n");
  return time_integration_scheme_;
}

std::string
Parser::GetModelMaterialParameters(int block_id) const
{
  auto it = model_blocks_.find(block_id);
  if (it == model_blocks_.end()) return "none";
  return material_strings_.at(it->second.material_key_);
}

int
Parser::GetBlockIdFromMaterial(const std::string& material_key) const
{
  for (auto const& kv : model_blocks_)
    if (kv.second.material_key_ == material_key) return kv.first;
  return -1;
}

std::string
Parser::GetOutputFieldString() const
{
  if (output_field_string_.empty())
    throw std::invalid_argument("\n**** Error in Parser::GetOutputFieldString(), output fields not found (possible input deck error?).");
  return output_field_string_;
}

}  // namespace nimble_b200
