// parse_vw.h -- reader of the .hr model files (eval/parse_vw.h:13, eval/parse_vw.cpp:8-30): one file per
// (target position, variant), lines "<name> <number>" with name = "Constant" or "<tagpos>_<variant>" and the
// number printed as a float ("-107.0") of which the integer prefix is kept (the reference's sscanf "%s %d").
#ifndef IDASH_B200_PARSE_VW_H
#define IDASH_B200_PARSE_VW_H

#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

// The reference's entry point: name -> coefficient; a repeated name keeps its last value.
// A file that cannot be opened prints "Cannot open file '<name>'" on stderr and aborts, like the reference.
std::unordered_map<std::string, int32_t> read(const std::string &file_name);

// Same parse, lines in file order (what read_model uses: no hash map per file).
std::vector<std::pair<std::string, int32_t>> read_lines(const std::string &file_name);

#endif
