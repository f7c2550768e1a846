// nimblesm_b200/host/exodus_output.h — Exodus II result writer with the surface of nimble::ExodusOutput
// (src/nimble_exodus_output.h, src/nimble_exodus_output.cc:62-396).  The ex_put_* calls are replaced by a
// direct NetCDF-3 (64-bit offset) writer producing the same dimensions / variables / names the reference's
// outputs contain, so that exodiff / epu / ParaView read the file like a reference one.
#pragma once
#include <exception>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace nimble_b200 {

class GenesisMesh;
namespace nc3 {
class Writer;
}

class ExodusOutput
{
 public:
  ExodusOutput();
  ~ExodusOutput();
  void
  Initialize(std::string const& filename, GenesisMesh const& genesis_mesh);
  std::string
  GetFileName() const
  {
    return filename_;
  }
  // Declares the variables and writes the mesh (src/nimble_exodus_output.cc:78-310).  Element variable names
  // are the alphabetically sorted union over blocks of the per-point and derived labels (std::set, :259-272).
  void
  InitializeDatabase(GenesisMesh const& genesis_mesh, std::vector<std::string> const& global_data_names,
                     std::vector<std::string> const& node_data_names,
                     std::map<int, std::vector<std::string>> const& elem_data_names,
                     std::map<int, std::vector<std::string>> const& derived_elem_data_names);
  // One time plane (src/nimble_exodus_output.cc:312-396): node_data[var][node]; elem_data[block][var][elem]
  void
  WriteStep(double time, std::vector<double> const& global_data, std::vector<std::vector<double>> const& node_data,
            std::map<int, std::vector<std::string>> const& elem_data_names,
            std::map<int, std::vector<std::vector<double>>> const& elem_data,
            std::map<int, std::vector<std::string>> const& derived_elem_data_names,
            std::map<int, std::vector<std::vector<double>>> const& derived_elem_data);
  // The same time plane, written by a background thread while the caller goes on stepping (the data is moved
  // into the job; at most one plane is in flight: the call first waits for the previous one).  At 8 M elements a
  // plane of 12 variables is 0.8 GB of file I/O, about as long as the 500 steps between two outputs.  Errors of
  // the writer thread surface at the next WriteStep / WriteStepAsync / Wait / Close.
  void
  WriteStepAsync(double time, std::vector<double> global_data, std::vector<std::vector<double>> node_data,
                 std::map<int, std::vector<std::string>> elem_data_names, std::map<int, std::vector<std::vector<double>>> elem_data,
                 std::map<int, std::vector<std::string>> derived_elem_data_names,
                 std::map<int, std::vector<std::vector<double>>> derived_elem_data);
  void
  Wait();
  int
  GetNumWrites() const
  {
    return exodus_write_count_;
  }
  void
  Close();

 private:
  void
  WritePlane(double time, std::vector<double> const& global_data, std::vector<std::vector<double>> const& node_data,
             std::map<int, std::vector<std::string>> const& elem_data_names,
             std::map<int, std::vector<std::vector<double>>> const& elem_data,
             std::map<int, std::vector<std::string>> const& derived_elem_data_names,
             std::map<int, std::vector<std::vector<double>>> const& derived_elem_data);
  std::string                  filename_;
  int                          dim_ = 3, num_nodes_ = 0, num_elements_ = 0, num_blocks_ = 0, num_global_blocks_ = 0;
  int                          num_node_sets_ = 0;
  std::vector<int>             block_ids_, all_block_ids_;
  std::map<int, int>           block_file_index_;  // block id -> 1-based index among ALL blocks of the file
  std::map<std::string, int>   elem_data_index_;   // variable name -> 1-based element variable index
  int                          num_node_vars_ = 0, num_global_vars_ = 0;
  int                          exodus_write_count_ = 0;
  std::unique_ptr<nc3::Writer> file_;
  std::thread                  writer_;
  std::exception_ptr           writer_error_;
};

}  // namespace nimble_b200
