// Host-side construction of the lane-centreline map table.
//
// This is the one piece of the reference's arithmetic that deliberately stays
// on the CPU: Map::Init (reference src/main.cpp:89-131) runs once per process
// over 181 rows and uses atan2/cos.  Running it through the host libm keeps
// the uploaded table bit-identical to the reference's, so every
// per-frame comparison downstream starts from the same numbers.  Compile
// without FMA contraction (-ffp-contract=off / nvcc -fmad=false).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "pp_internal.h"

namespace ppi {

namespace {
inline int wrap(int i, int n) { return ((i % n) + n) % n; }  // get_waypoint, :134-137 (|i| < n here)
}  // namespace

int build_map_table(const double *wx, const double *wy, int n, std::vector<double> &table) {
  if (!wx || !wy) return PP_E_ARG;
  if (n < 2 || n > 65536) return PP_E_RANGE;
  table.assign((size_t)n * PP_MAP_STRIDE, 0.0);
  auto row = [&](int i) { return table.data() + (size_t)wrap(i, n) * PP_MAP_STRIDE; };

  // reference line, :94-99
  for (int i = 0; i < n; i++) {
    row(i)[0] = wx[i];
    row(i)[1] = wy[i];
  }
  // unit normal of the segment arriving at waypoint i, :101-109
  for (int i = 0; i < n; i++) {
    double *w = row(i);
    const double *p = row(i - 1);
    const double ddx = w[0] - p[0];
    const double ddy = w[1] - p[1];
    const double dl = std::sqrt(ddx * ddx + ddy * ddy);
    w[8] = ddy / dl;
    w[9] = -ddx / dl;
  }
  // lane centres on the averaged normal, stretched by 1/cos(alpha), :111-130
  for (int i = 0; i < n; i++) {
    double *w = row(i);
    const double *q = row(i + 1);
    double ax = (w[8] + q[8]) / 2;
    double ay = (w[9] + q[9]) / 2;
    const double a_seg = std::atan2(w[9], w[8]);
    const double a_avg = std::atan2(ay, ax);
    const double c = std::cos(a_avg - a_seg);
    ax /= c;
    ay /= c;
    for (int lane = 0; lane < PP_NUM_LANES; lane++) {
      const double lane_width = 4.0;  // :86
      const double off = lane_width * (lane + 0.5);
      w[2 + 2 * lane] = w[0] + ax * off;
      w[3 + 2 * lane] = w[1] + ay * off;
    }
  }
  // get_lane_length(i, lane), :138-142, precomputed with the same expression
  for (int i = 0; i < n; i++) {
    double *w = row(i);
    const double *p = row(i - 1);
    for (int lane = 0; lane < PP_NUM_LANES; lane++) {
      const double ex = w[2 + 2 * lane] - p[2 + 2 * lane];
      const double ey = w[3 + 2 * lane] - p[3 + 2 * lane];
      w[10 + lane] = std::sqrt(ex * ex + ey * ey);
    }
  }
  return PP_OK;
}

int read_map_csv(const char *path, std::vector<double> &wx, std::vector<double> &wy) {
  if (!path) return PP_E_ARG;
  FILE *fp = std::fopen(path, "r");
  if (!fp) return PP_E_IO;
  char line[1024];
  wx.clear();
  wy.clear();
  // The reference reads "x y s dx dy" with operator>> and keeps x, y as
  // double (:1176-1187); s, dx, dy go through float and are never used.
  while (std::fgets(line, sizeof line, fp)) {
    char *end1 = nullptr, *end2 = nullptr;
    const double x = std::strtod(line, &end1);
    if (end1 == line) continue;
    const double y = std::strtod(end1, &end2);
    if (end2 == end1) continue;
    wx.push_back(x);
    wy.push_back(y);
  }
  std::fclose(fp);
  return wx.size() >= 2 ? PP_OK : PP_E_IO;
}

}  // namespace ppi
