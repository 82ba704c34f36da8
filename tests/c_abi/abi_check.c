/* Plain C client of include/fclgpu.h: proves the header is C (no C++/torch types), the library
 * links from C, the host-side BVH build works without a GPU and compute calls fail loudly (no CPU
 * fallback) when no device is present. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fclgpu.h"

int main(void) {
  /* unit box, 12 triangles (the reference's test_fcl_bvh_models.cpp builds the same shape) */
  const double v[8 * 3] = {-1, -1, -1, 1, -1, -1, 1, 1, -1, -1, 1, -1, -1, -1, 1, 1, -1, 1, 1, 1, 1, -1, 1, 1};
  const int32_t t[12 * 3] = {0, 2, 1, 0, 3, 2, 4, 5, 6, 4, 6, 7, 0, 1, 5, 0, 5, 4, 2, 3, 7, 2, 7, 6, 1, 2, 6, 1, 6, 5, 0, 4, 7, 0, 7, 3};
  fclgpu_bvh* bvh = NULL;
  int rc = fclgpu_bvh_build_obbrss(v, 8, t, 12, FCLGPU_SPLIT_METHOD_MEAN, &bvh);
  if (rc != FCLGPU_OK || !bvh) { printf("build failed %d\n", rc); return 1; }
  if (fclgpu_bvh_num_nodes(bvh) != 23 || fclgpu_bvh_num_tris(bvh) != 12) { printf("bad counts\n"); return 1; }
  int32_t fc[23];
  double ext[23 * 3];
  fclgpu_bvh_get(bvh, fc, NULL, NULL, ext, NULL, NULL, NULL, NULL);
  int leaves = 0;
  for (int i = 0; i < 23; ++i) leaves += fc[i] < 0;
  if (leaves != 12) { printf("bad leaves %d\n", leaves); return 1; }
  fclgpu_bvh* none = NULL;
  if (fclgpu_bvh_build_obbrss(v, 8, t, 0, 0, &none) != FCLGPU_ERR_BUILD_EMPTY_MODEL || none != NULL) { printf("empty model not rejected\n"); return 1; }
  if (fclgpu_abi_version() != FCLGPU_ABI_VERSION) return 1;
  if (sizeof(fclgpu_contact) != 64) { printf("contact size %zu\n", sizeof(fclgpu_contact)); return 1; }
  double m16[16] = {0, 1, 0, 0, -1, 0, 0, 0, 0, 0, 1, 0, 5, 6, 7, 1}, pose[12];
  fclgpu_pose_from_colmajor4x4(m16, pose);
  if (pose[1] != -1 || pose[3] != 1 || pose[9] != 5 || pose[11] != 7) { printf("pose conversion\n"); return 1; }
  fclgpu_model* m = NULL;
  rc = fclgpu_model_from_bvh(0, bvh, &m);
  if (fclgpu_device_count() == 0) {
    if (rc != FCLGPU_ERR_NO_DEVICE) { printf("expected NO_DEVICE, got %d (%s)\n", rc, fclgpu_last_error()); return 1; }
    printf("no device: upload refused with %d (%s)\n", rc, fclgpu_last_error());
  } else {
    if (rc != FCLGPU_OK) { printf("upload failed %d (%s)\n", rc, fclgpu_last_error()); return 1; }
    fclgpu_collision_request req = {1, 0, 0, 0, FCLGPU_CONTACT_FULL, 0};
    double tf[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5, 0, 0};
    int32_t n = -1;
    rc = fclgpu_collide_batch_host(m, m, 1, tf, NULL, &req, &n, NULL, 0, NULL, NULL, NULL);
    if (rc != FCLGPU_OK || n != 1) { printf("collide failed rc=%d n=%d\n", rc, n); return 1; }
    /* mesh <-> sphere: a unit-radius sphere 3 away from the box centre along x is 1 away from the face x = 1 */
    fclgpu_distance_request dreq = {1, 0, 0.0, 0.0};
    double stf[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 3.0, 0.25, -0.5}, d = 0, p1[3], p2[3];
    int32_t b1 = -5, b2 = -5;
    rc = fclgpu_distance_mesh_sphere_batch_host(m, 1.0, 1, NULL, stf, &dreq, &d, p1, p2, &b1, &b2, NULL, NULL);
    if (rc != FCLGPU_OK || d != 1.0 || b2 != -1 || b1 != 8 || p1[0] != 1.0 || p1[2] != -0.5 || p2[0] != -1.0) {
      printf("sphere distance failed rc=%d d=%.17g b1=%d b2=%d p1x=%g p2x=%g\n", rc, d, b1, b2, p1[0], p2[0]);
      return 1;
    }
    fclgpu_model_destroy(m);
    printf("device run ok\n");
  }
  /* argument checks need no device */
  {
    fclgpu_distance_request dreq = {1, 0, 0.0, 0.0};
    double d;
    if (fclgpu_distance_mesh_sphere_batch_host(NULL, 1.0, 1, NULL, NULL, &dreq, &d, NULL, NULL, NULL, NULL, NULL, NULL) != FCLGPU_ERR_INVALID_ARGUMENT ||
        fclgpu_distance_mesh_sphere_batch(NULL, -1.0, 1, NULL, NULL, &dreq, &d, NULL, NULL, NULL, NULL, NULL, NULL, NULL) != FCLGPU_ERR_INVALID_ARGUMENT) {
      printf("sphere distance argument checks\n");
      return 1;
    }
  }
  fclgpu_bvh_destroy(bvh);
  printf("abi_check ok\n");
  return 0;
}
