// mesh_io.cpp — Wavefront OBJ reader / writer of the reference as product code:
// loadOBJFile / saveOBJFile, test/test_fcl_utility.h:194-280, 283-309 -- the on-disk format of the benchmark's meshes
// (test/fcl_resources/env.obj, rob.obj).  Behaviour kept, quirks included, so that the same file yields the same
// vertices and triangles as in the reference's own tests:
//   * a line is split at blanks / tabs / CR / LF; lines whose first token starts with '#' are skipped, and so is every
//     line whose first token starts with anything but 'v' or 'f' (the fixtures' first line "6540 2180" is one of those);
//   * "vn ..." and "vt ..." are not stored but REMEMBERED (has_normal / has_texture), every other 'v*' token is a vertex
//     with three atof() coordinates;
//   * "f a b c ...": indices are 1-based, a corner "v/vt/vn" is read with atoi (up to the first '/').  Files WITH
//     normals or textures are fanned (0, t+1, t+2); files with neither emit the triangle (0, 1, 2) once per fan step
//     (the reference's no-normal branch ignores t, :250-255) -- identical for triangle-only files such as the fixtures.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/fclgpu.h"

extern "C" int fclgpu_load_obj(const char* path, double** vertices, int32_t* num_vertices, int32_t** triangles, int32_t* num_tris) {
  if (!path || !vertices || !num_vertices || !triangles || !num_tris) return FCLGPU_ERR_INVALID_ARGUMENT;
  *vertices = nullptr;
  *triangles = nullptr;
  *num_vertices = *num_tris = 0;
  FILE* file = std::fopen(path, "rb");
  if (!file) return FCLGPU_ERR_INCORRECT_DATA;  // the reference prints "file not exist" and returns empty arrays
  std::vector<double> pts;
  std::vector<int32_t> tris;
  bool has_normal = false, has_texture = false;
  char line[2000];
  while (std::fgets(line, sizeof line, file)) {
    char* first = std::strtok(line, "\r\n\t ");
    if (!first || first[0] == '#' || first[0] == 0) continue;
    if (first[0] == 'v') {
      if (first[1] == 'n') {
        has_normal = true;
      } else if (first[1] == 't') {
        has_texture = true;
      } else {
        for (int k = 0; k < 3; ++k) {
          const char* s = std::strtok(nullptr, "\t ");
          pts.push_back(s ? std::atof(s) : 0.0);  // (the reference would dereference NULL on a short line)
        }
      }
    } else if (first[0] == 'f') {
      const char* data[30];
      int n = 0;
      while (n < 30 && (data[n] = std::strtok(nullptr, "\t \r\n")) != nullptr)
        if (std::strlen(data[n])) n++;
      for (int t = 0; t < n - 2; ++t) {
        if (!has_texture && !has_normal) {
          for (int i = 0; i < 3; ++i) tris.push_back(std::atoi(data[i]) - 1);
        } else {
          tris.push_back(std::atoi(data[0]) - 1);
          tris.push_back(std::atoi(data[t + 1]) - 1);
          tris.push_back(std::atoi(data[t + 2]) - 1);
        }
      }
    }
  }
  std::fclose(file);
  *num_vertices = (int32_t)(pts.size() / 3);
  *num_tris = (int32_t)(tris.size() / 3);
  if (!pts.empty()) {
    *vertices = (double*)std::malloc(pts.size() * sizeof(double));
    if (!*vertices) return FCLGPU_ERR_MODEL_OUT_OF_MEMORY;
    std::memcpy(*vertices, pts.data(), pts.size() * sizeof(double));
  }
  if (!tris.empty()) {
    *triangles = (int32_t*)std::malloc(tris.size() * sizeof(int32_t));
    if (!*triangles) {
      std::free(*vertices);
      *vertices = nullptr;
      return FCLGPU_ERR_MODEL_OUT_OF_MEMORY;
    }
    std::memcpy(*triangles, tris.data(), tris.size() * sizeof(int32_t));
  }
  return FCLGPU_OK;
}

// saveOBJFile, test/test_fcl_utility.h:283-309: "v x y z" per vertex, "f a b c" (1-based) per triangle.  Coordinates are
// written with 17 significant digits so that a save / load round trip is exact (the reference streams them with the
// default precision of 6).
extern "C" int fclgpu_save_obj(const char* path, const double* vertices, int32_t num_vertices, const int32_t* triangles,
                               int32_t num_tris) {
  if (!path || (num_vertices > 0 && !vertices) || (num_tris > 0 && !triangles)) return FCLGPU_ERR_INVALID_ARGUMENT;
  FILE* f = std::fopen(path, "wb");
  if (!f) return FCLGPU_ERR_INCORRECT_DATA;
  for (int32_t i = 0; i < num_vertices; ++i)
    std::fprintf(f, "v %.17g %.17g %.17g\n", vertices[3 * (size_t)i], vertices[3 * (size_t)i + 1], vertices[3 * (size_t)i + 2]);
  for (int32_t i = 0; i < num_tris; ++i)
    std::fprintf(f, "f %d %d %d\n", triangles[3 * (size_t)i] + 1, triangles[3 * (size_t)i + 1] + 1, triangles[3 * (size_t)i + 2] + 1);
  return std::fclose(f) == 0 ? FCLGPU_OK : FCLGPU_ERR_UNKNOWN;
}

extern "C" void fclgpu_free(void* p) { std::free(p); }
