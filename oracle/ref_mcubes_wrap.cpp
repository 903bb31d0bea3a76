// TEST INFRASTRUCTURE ONLY.  extern "C" face of the reference's own marching cubes, compiled from the sources where
// they lie (/root/reference/ONet/im2mesh/utils/libmcubes/marchingcubes.{h,cpp}) into oracle/_ref/ by oracle/Makefile.
// It restates only the 10 lines of pywrapper.cpp:90-107 (`marching_cubes(arr, isovalue)`: lower = 0, upper = shape - 1,
// the volume read through an (int x, int y, int z) accessor), because the Python wrapper itself does not build
// against numpy 2 (PyArray_DOUBLE).  Nothing from the reference is copied into this repository.
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "marchingcubes.h"   // -I/root/reference/ONet/im2mesh/utils/libmcubes

namespace {
struct Volume {
  const double* v;
  long ny, nz;
  double operator()(int x, int y, int z) const { return v[((long)x * ny + y) * nz + z]; }
};
std::vector<double> g_verts;
std::vector<size_t> g_faces;
}  // namespace

extern "C" {
// runs the reference; returns the counts; the arrays are then fetched with refmc_fetch
void refmc_run(const double* volume, long nx, long ny, long nz, double isovalue, long* n_vert_doubles, long* n_indices) {
  g_verts.clear();
  g_faces.clear();
  double lower[3] = {0, 0, 0};
  double upper[3] = {(double)(nx - 1), (double)(ny - 1), (double)(nz - 1)};
  Volume f{volume, ny, nz};
  mc::marching_cubes<double>(lower, upper, (int)nx, (int)ny, (int)nz, f, isovalue, g_verts, g_faces);
  *n_vert_doubles = (long)g_verts.size();
  *n_indices = (long)g_faces.size();
}
void refmc_fetch(double* verts, long long* indices) {
  if (!g_verts.empty()) std::memcpy(verts, g_verts.data(), g_verts.size() * sizeof(double));
  for (size_t i = 0; i < g_faces.size(); ++i) indices[i] = (long long)g_faces[i];
}
}
