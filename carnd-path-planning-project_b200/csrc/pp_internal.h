// Internal declarations shared by the translation units of libpp_b200.so.
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/pp.h"

// rows of wrap-around padding on either side of the device copy of the table
#define PPD_PAD_ROWS 24

struct pp_map {
  int n = 0;                  // waypoints
  std::vector<double> table;  // n * PP_MAP_STRIDE, host copy (row layout of pp.h)
  double *dev_table = nullptr;  // device copy, padded (nullptr if no CUDA device was usable)
  int device = -1;              // CUDA device the table lives on
};

namespace ppi {

// Host-side Map::Init (reference src/main.cpp:89-131): fills `table`.
int build_map_table(const double *wx, const double *wy, int n, std::vector<double> &table);
// CSV reader with the reference's parsing (src/main.cpp:1171-1191).
int read_map_csv(const char *path, std::vector<double> &wx, std::vector<double> &wy);

// Upload (pp_api.cu).  Returns PP_OK or PP_E_CUDA.
int upload_map(pp_map *m);
void free_map_device(pp_map *m);

void set_cuda_error(const char *what, int cuda_err, const char *text);
void count_launch(int n = 1);

}  // namespace ppi
