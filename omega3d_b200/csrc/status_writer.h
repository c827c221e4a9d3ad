// status_writer.h - the line-per-step status file of the reference (host I/O, no kernel).
//
// Written from the behaviour of StatusFile (src/StatusFile.h:31-76, src/StatusFile.cpp:38-137) and of
// Simulation::dump_stats_to_status / calculate_simple_forces (src/Simulation.cpp:851-924): a data set starts with a header
// line of the value names ("# " in front for the space-separated .dat flavour, bare and comma-separated for .csv), data sets
// after the first are preceded by an empty line, every value goes through operator<< of a std::ofstream opened in append
// mode (so floats print as %g with 6 significant digits, ints as %d) - the files are byte-identical to the reference's for
// the same values (tests/golden/status.npz, tests/test_status.py).
#pragma once
#include <fstream>
#include <string>
#include <variant>
#include <vector>

struct o3d_status {
  bool csv = false;
  int num_sims = 0;      // data sets written to this file so far
  int num_lines = 0;     // data lines of the current set
  std::string fn;
  std::vector<std::string> names;
  std::vector<std::variant<float, int>> vals;
  // calculate_simple_forces keeps these as function statics (src/Simulation.cpp:903-904)
  double last_time = 0.0;
  float last_impulse[3] = {0.0f, 0.0f, 0.0f};

  void reset_sim() { num_lines = 0; }
  void append(const std::string& name, float v) { names.push_back(name); vals.emplace_back(v); }
  void append(const std::string& name, int v) { names.push_back(name); vals.emplace_back(v); }
  bool write_line() {
    std::ofstream out(fn, std::ios::app);
    if (!out) { vals.clear(); return false; }
    const char* sep = csv ? "," : " ";
    if (num_lines == 0) {
      if (num_sims > 0) out << std::endl;
      if (!csv) out << "# ";
      for (size_t i = 0; i < names.size(); ++i) {
        out << names[i];
        if (i + 1 < names.size()) out << sep;
      }
      out << std::endl;
      ++num_sims;
    }
    for (size_t i = 0; i < vals.size(); ++i) {
      std::visit([&out](const auto& v) { out << v; }, vals[i]);
      if (i + 1 < vals.size()) out << sep;
      else out << std::endl;
    }
    ++num_lines;
    vals.clear();      // the reference clears the values but NOT the names: a header written by a later data set repeats them
    return (bool)out;
  }
};
