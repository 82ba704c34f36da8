/* fclgpu.h — C ABI of the B200-native batched BVHModel<OBBRSS<double>> mesh-mesh
 * collide()/distance() path.
 *
 * This is the drop-in boundary for ONE cell of the reference's dispatch tables:
 *   collision_matrix[BV_OBBRSS][BV_OBBRSS]  (include/fcl/narrowphase/detail/collision_func_matrix-inl.h:851)
 *   distance_matrix [BV_OBBRSS][BV_OBBRSS]  (include/fcl/narrowphase/detail/distance_func_matrix-inl.h:663)
 * evaluated over a batch of pose pairs.  All citations are relative to the
 * reference tree (flexible-collision-library/fcl).
 *
 * Conventions
 *  - plain C types, caller-owned buffers, no exceptions; every entry point
 *    returns an int status (0 = ok, negative = error, see fclgpu_status).
 *  - a pose ("tf") is 12 doubles: rotation R row-major (9) then translation t (3),
 *    p_world = R p + t.  (The reference's Transform3d is a 4x4 column-major
 *    Eigen::Isometry, include/fcl/common/types.h:70-92; fclgpu_pose_from_colmajor4x4
 *    converts.)  A NULL tf array means identity for every query.
 *  - *_batch entry points take DEVICE pointers (memory resident on the model's
 *    device) and enqueue on `stream` (a cudaStream_t passed as void*; NULL = the
 *    legacy default stream) without synchronising.
 *  - *_batch_host entry points take HOST pointers, stage through pinned memory,
 *    copy, launch, copy back and return when the results are in the caller's
 *    buffers (this is what a single fcl::collide()/fcl::distance() call maps to).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with FCLGPU_ERR_NO_DEVICE.
 */
#ifndef FCLGPU_H_
#define FCLGPU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCLGPU_ABI_VERSION 3 /* 2: contact_offsets[i] = start of query i's block (order of the blocks: see below);
                               * 3: fclgpu_collision_request::contact_format (compact contact records) */

/* Status codes.  -1..-8 mirror the reference's BVHReturnCode
 * (include/fcl/geometry/bvh/BVH_internal.h:61-72). */
typedef enum fclgpu_status {
  FCLGPU_OK = 0,
  FCLGPU_ERR_MODEL_OUT_OF_MEMORY = -1,
  FCLGPU_ERR_BUILD_OUT_OF_SEQUENCE = -2,
  FCLGPU_ERR_BUILD_EMPTY_MODEL = -3,
  FCLGPU_ERR_BUILD_EMPTY_PREVIOUS_FRAME = -4,
  FCLGPU_ERR_UNSUPPORTED_FUNCTION = -5,
  FCLGPU_ERR_UNUPDATED_MODEL = -6,
  FCLGPU_ERR_INCORRECT_DATA = -7,
  FCLGPU_ERR_UNKNOWN = -8,
  FCLGPU_ERR_INVALID_ARGUMENT = -20,
  FCLGPU_ERR_NO_DEVICE = -21,
  FCLGPU_ERR_CUDA = -22,
  FCLGPU_ERR_CONTACT_OVERFLOW = -23, /* per-pose or pool capacity too small; counts are still exact */
  FCLGPU_ERR_STACK_OVERFLOW = -24,   /* traversal stack exhausted (tree deeper than supported) */
  FCLGPU_ERR_INPUT_STALLED = -25,    /* host API: a pose chunk did not reach the device within the kernel's wait budget */
  FCLGPU_ERR_COMM = -26              /* fclgpu_comm_*: NCCL reported an error (fclgpu_comm_last_error) */
} fclgpu_status;

/* Split rules of the reference's BVSplitter (include/fcl/geometry/bvh/detail/BV_splitter.h). */
typedef enum fclgpu_split_method {
  FCLGPU_SPLIT_METHOD_MEAN = 0,
  FCLGPU_SPLIT_METHOD_MEDIAN = 1,
  FCLGPU_SPLIT_METHOD_BV_CENTER = 2
} fclgpu_split_method;

/* One contact; field meaning as fcl::Contact<double> (include/fcl/narrowphase/contact.h:48-91)
 * with b1/b2 = triangle ids in model1/model2.  In binary mode (enable_contact == 0) only
 * b1/b2 are defined, as in the reference.  64 bytes. */
typedef struct fclgpu_contact {
  int32_t b1, b2;
  double normal[3];
  double pos[3];
  double penetration_depth;
} fclgpu_contact;

/* Compact contact records (not in the reference; fclgpu_collision_request::contact_format).  A contact list of the
 * full 64-byte records is 2 GB per million queries at 32 contacts each, and the copy to the host then takes twice as
 * long as the traversal; a caller that only needs WHICH triangles touch, or single-precision geometry, asks for these:
 *   FCLGPU_CONTACT_IDS  8 bytes: the two primitive ids;
 *   FCLGPU_CONTACT_F32 40 bytes: ids + normal, position and depth rounded to nearest float from the same FP64 values.
 * Same lists, same order, same truncation as the full records.  Supported by fclgpu_collide_batch[_host] on its default
 * (one-launch) contact path; elsewhere a format other than FULL returns FCLGPU_ERR_UNSUPPORTED_FUNCTION. */
enum { FCLGPU_CONTACT_FULL = 0, FCLGPU_CONTACT_IDS = 1, FCLGPU_CONTACT_F32 = 2 };
typedef struct fclgpu_contact_ids {
  int32_t b1, b2;
} fclgpu_contact_ids;
typedef struct fclgpu_contact_f32 {
  int32_t b1, b2;
  float normal[3];
  float pos[3];
  float penetration_depth;
  float reserved;
} fclgpu_contact_f32;

/* Honoured fields of fcl::CollisionRequest (include/fcl/narrowphase/collision_request.h:52-106).
 * enable_cost must be 0 (cost sources are not on this path: FCLGPU_ERR_UNSUPPORTED_FUNCTION). */
typedef struct fclgpu_collision_request {
  int64_t num_max_contacts; /* default 1; 0 -> every query returns 0 contacts (collision-inl.h:111-115) */
  int32_t enable_contact;   /* default 0 */
  int32_t enable_cost;      /* must be 0 */
  /* Not an fcl::CollisionRequest field: contact slots staged per query before they reach the pool when a contact
   * list is requested.  0 = library default (option "contact_stride", 1024).  A query with more contacts than
   * min(num_max_contacts, stage_capacity) reports FCLGPU_ERR_CONTACT_OVERFLOW (counts stay exact): rerun with
   * stage_capacity = max(num_contacts).  Per call, so concurrent callers do not share a setting. */
  int64_t stage_capacity;
  /* Not an fcl::CollisionRequest field either: layout of the records written to `contacts` (FCLGPU_CONTACT_*; 0 = the
   * full fclgpu_contact).  `contacts` then points to contact_capacity records of that format. */
  int32_t contact_format;
  int32_t reserved;
} fclgpu_collision_request;

/* Honoured fields of fcl::DistanceRequest (include/fcl/narrowphase/distance_request.h:52-113).
 * rel_err / abs_err are accepted for source compatibility and ignored, exactly as the
 * reference does on this path (mesh_distance_traversal_node-inl.h:96-105 vs :606-633). */
typedef struct fclgpu_distance_request {
  int32_t enable_nearest_points;
  int32_t enable_signed_distance; /* no effect for meshes (triDistance >= 0) */
  double rel_err, abs_err;        /* ignored */
} fclgpu_distance_request;

/* ---------------------------------------------------------------------------------------
 * Host-side BVH construction: BVHModel<OBBRSS<double>>::beginModel/addSubModel/endModel
 * (include/fcl/geometry/bvh/BVH_model-inl.h:207-253, 383-517, 833-938) with
 * FitImpl<OBBRSS> (detail/BV_fitter-inl.h:449-477) and BVSplitter (detail/BV_splitter-inl.h).
 * The result is the flattened node tree the upload step consumes.
 * ------------------------------------------------------------------------------------- */
typedef struct fclgpu_bvh fclgpu_bvh; /* host object */

int fclgpu_bvh_build_obbrss(const double* vertices /* nv x 3 */, int32_t num_vertices,
                            const int32_t* triangles /* nt x 3 */, int32_t num_tris,
                            int32_t split_method, fclgpu_bvh** out);
void fclgpu_bvh_destroy(fclgpu_bvh* bvh);
int32_t fclgpu_bvh_num_nodes(const fclgpu_bvh* bvh);
int32_t fclgpu_bvh_num_tris(const fclgpu_bvh* bvh);
/* Copies the node arrays out (any pointer may be NULL).  axis9 is row-major: axis9[3*r+c],
 * column c = c-th box axis.  tri_verts9 = de-indexed triangle vertices (p1 p2 p3). */
int fclgpu_bvh_get(const fclgpu_bvh* bvh, int32_t* first_child, double* axis9, double* obb_To3,
                   double* obb_extent3, double* rss_To3, double* rss_l2, double* rss_r,
                   double* tri_verts9);

/* Top-down refit on the host: BVHModel::beginReplaceModel / replaceSubModel / endReplaceModel(
 * refit = true, bottomup = false) (BVH_model-inl.h:521-620, refitTree_topdown :1064-1076): new vertex
 * positions (same count), same tree, every BV refitted over its stored primitive range. */
int fclgpu_bvh_refit_topdown(fclgpu_bvh* bvh, const double* vertices, int32_t num_vertices);
/* Bottom-up refit on the host: endReplaceModel(refit = true, bottomup = true), the reference's DEFAULT
 * (BVH_model.h:128, BVH_model-inl.h:952-1037): a leaf gets the closed-form fit of its triangle (fit3,
 * math/bv/utility-inl.h:92-117, 208-230), an inner node the merge of its children's volumes (OBB::operator+,
 * OBB-inl.h:161-369; RSS::operator+, RSS-inl.h:313-371).  The reference's quirks are kept bit for bit (the `else if`
 * of merge_smalldist, the transposed eigenvector matrix and the stale third axis of RSS::operator+), so -- as in
 * the reference -- a volume refitted this way need not contain its subtree.  Afterwards the RSS of a node no longer
 * shares the OBB's axes (fclgpu_bvh_get_rss_axis). */
int fclgpu_bvh_refit_bottomup(fclgpu_bvh* bvh, const double* vertices, int32_t num_vertices);
/* rss.axis per node (row-major 9).  Returns 1 when they are separate from the OBB's (after a bottom-up refit), 0 when
 * shared (the array then receives the OBB's axes); < 0 on error. */
int fclgpu_bvh_get_rss_axis(const fclgpu_bvh* bvh, double* rss_axis9);
int32_t fclgpu_bvh_num_vertices(const fclgpu_bvh* bvh);
/* BVNodeBase::first_primitive / num_primitives per node, BVHModel::primitive_indices, tri_indices
 * (any pointer may be NULL). */
int fclgpu_bvh_get_partition(const fclgpu_bvh* bvh, int32_t* first_primitive, int32_t* num_primitives,
                             int32_t* primitive_indices, int32_t* tri_indices3);

/* Mesh files: the reference's Wavefront OBJ reader / writer, loadOBJFile / saveOBJFile
 * (test/test_fcl_utility.h:194-309), the format of test/fcl_resources/env.obj and rob.obj.  fclgpu_load_obj allocates
 * *vertices (num_vertices x 3) and *triangles (num_tris x 3, 0-based) with malloc: release them with fclgpu_free.
 * A missing file returns FCLGPU_ERR_INCORRECT_DATA (the reference prints "file not exist" and leaves the arrays empty). */
int fclgpu_load_obj(const char* path, double** vertices, int32_t* num_vertices, int32_t** triangles, int32_t* num_tris);
int fclgpu_save_obj(const char* path, const double* vertices, int32_t num_vertices, const int32_t* triangles, int32_t num_tris);
void fclgpu_free(void* p);

/* ---------------------------------------------------------------------------------------
 * Upload: flattens BVNode<OBBRSS<double>>[] (include/fcl/geometry/bvh/BV_node.h:50-72,
 * BVH_model.h:160-203) + triangles into device records ("HBM layout").
 * The arrays are exactly what a BVHModel<OBBRSS<double>> holds: getBV(i).first_child,
 * .bv.obb.{axis,To,extent}, .bv.rss.{To,l,r} (rss.axis == obb.axis by construction,
 * BV_fitter-inl.h:464) and vertices[tri_indices[t][k]].
 * ------------------------------------------------------------------------------------- */
typedef struct fclgpu_model fclgpu_model; /* device-resident model */

int fclgpu_model_create_obbrss(int device, int32_t n_nodes, const int32_t* first_child,
                               const double* axis9, const double* obb_To3, const double* obb_extent3,
                               const double* rss_To3, const double* rss_l2, const double* rss_r,
                               int32_t n_tris, const double* tri_verts9, fclgpu_model** out);
/* The same with separate RSS axes (rss_axis9 == NULL: shared).  endModel() and the top-down refit give both volumes
 * the same axes; the reference's BOTTOM-UP refit (BVHModel::refitTree_bottomup, BVH_model-inl.h:952-1025) merges the
 * children's OBBs and RSSs separately (OBBRSS::operator+, OBBRSS-inl.h:95-101), after which they differ. */
int fclgpu_model_create_obbrss2(int device, int32_t n_nodes, const int32_t* first_child, const double* axis9,
                                const double* obb_To3, const double* obb_extent3, const double* rss_axis9,
                                const double* rss_To3, const double* rss_l2, const double* rss_r, int32_t n_tris,
                                const double* tri_verts9, fclgpu_model** out);
int fclgpu_model_from_bvh(int device, const fclgpu_bvh* bvh, fclgpu_model** out);
/* On-device build: BVHModel::endModel() for BVH_MODEL_TRIANGLES + OBBRSS
 * (BVH_model-inl.h:450-517, 833-938) executed level by level on the GPU from the host arrays
 * `vertices` (num_vertices x 3) and `triangles` (num_tris x 3).  Same tree, node numbering,
 * primitive order and volumes as fclgpu_bvh_build_obbrss + fclgpu_model_from_bvh, bit for bit.
 * split_method: any of the three rules; the median rule (BV_splitter-inl.h:603-657) finds the one or two middle
 * projections by a cooperative bitwise selection instead of sorting.  Synchronous. */
int fclgpu_model_build_obbrss(int device, const double* vertices, int32_t num_vertices,
                              const int32_t* triangles, int32_t num_tris, int32_t split_method,
                              fclgpu_model** out);
/* Topology of a device model: first_child per node and, when the model carries it, the build
 * partition (BVNodeBase::first_primitive / num_primitives, BVHModel::primitive_indices). Any pointer may be NULL. */
int fclgpu_model_get_topology(const fclgpu_model* m, int32_t* first_child, int32_t* first_primitive,
                              int32_t* num_primitives, int32_t* primitive_indices);
int fclgpu_model_destroy(fclgpu_model* m);
/* Refit topology for a model created from raw node arrays (fclgpu_model_from_bvh sets it itself). */
int fclgpu_model_set_partition(fclgpu_model* m, int32_t num_vertices, const int32_t* tri_indices3,
                               const int32_t* first_primitive, const int32_t* num_primitives,
                               const int32_t* primitive_indices);
/* On-device top-down refit: same semantics as fclgpu_bvh_refit_topdown, but the
 * node records in HBM are recomputed by a kernel (one thread per node, sums in the reference's
 * order => bit-identical BVs).  `vertices` (num_vertices x 3) is a device pointer when
 * vertices_on_device != 0, else a host pointer.  Asynchronous on `stream`. */
int fclgpu_model_refit_topdown(fclgpu_model* m, const double* vertices, int32_t num_vertices,
                               int32_t vertices_on_device, void* stream);
/* On-device bottom-up refit: same semantics and bits as fclgpu_bvh_refit_bottomup, one launch per tree height (every
 * node of a height in parallel).  Needs the triangle indices (fclgpu_model_set_partition / fclgpu_model_from_bvh /
 * fclgpu_model_build_obbrss).  Asynchronous on `stream`. */
int fclgpu_model_refit_bottomup(fclgpu_model* m, const double* vertices, int32_t num_vertices,
                                int32_t vertices_on_device, void* stream);
/* rss.axis of every node as stored in HBM (row-major 9 per node). */
int fclgpu_model_download_rss_axis(const fclgpu_model* m, double* rss_axis9);
/* Copies the FP64 node records back to the host (testing / inspection; any pointer may be NULL). */
int fclgpu_model_download(const fclgpu_model* m, double* axis9, double* obb_To3, double* obb_extent3,
                          double* rss_To3, double* rss_l2, double* rss_r, double* tri_verts9);
int32_t fclgpu_model_num_nodes(const fclgpu_model* m);
int32_t fclgpu_model_num_tris(const fclgpu_model* m);
int fclgpu_model_device(const fclgpu_model* m);

/* ---------------------------------------------------------------------------------------
 * Batched collide: query i evaluates fcl::collide(m1, tf1[i], m2, tf2[i], request, result_i)
 * with a fresh result_i (collision-inl.h:95-207 -> orientedMeshCollide,
 * collision_func_matrix-inl.h:571-590).
 *   num_contacts[i]  = result_i.numContacts()
 *   contacts         : dense pool; query i's contacts, in the reference's DFS order, occupy the contiguous block
 *                      contacts[contact_offsets[i] .. contact_offsets[i] + num_contacts[i]).  Blocks are disjoint and
 *                      fill [0, contact_offsets[n]) without gaps.  In the reference every query has its own
 *                      CollisionResult, so the ORDER OF THE BLOCKS is not reference semantics: by default (option
 *                      "contact_order" = 0) each block is appended when its query retires on the GPU -- one kernel
 *                      launch, one DRAM write per contact; with "contact_order" = 1 blocks come in query order and
 *                      contact_offsets is the exclusive prefix sum of num_contacts (scan + compaction passes).
 *   contact_offsets  : n+1 entries (start of each query's block; [n] = total number of contacts); may be NULL
 *                      together with contacts when only counts/verdicts are wanted.
 *   contact_capacity : slots available in `contacts`.
 *   n_bv / n_leaf    : optional per-query counters (num_bv_tests / num_leaf_tests of the
 *                      reference's traversal node, bvh_collision_traversal_node.h:92-94).
 * ------------------------------------------------------------------------------------- */
int fclgpu_collide_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n,
                         const double* tf1, const double* tf2,
                         const fclgpu_collision_request* request, int32_t* num_contacts,
                         fclgpu_contact* contacts, int64_t contact_capacity,
                         int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf, void* stream);

int fclgpu_collide_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n,
                              const double* tf1, const double* tf2,
                              const fclgpu_collision_request* request, int32_t* num_contacts,
                              fclgpu_contact* contacts, int64_t contact_capacity,
                              int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf);

/* ---------------------------------------------------------------------------------------
 * Batched continuous collision: query i evaluates
 *   fcl::continuousCollide(m1, tf1_beg[i], tf1_end[i], m2, tf2_beg[i], tf2_end[i], request, result_i)
 * (narrowphase/continuous_collision-inl.h:441-452) for request.ccd_motion_type = CCDM_TRANS (both bodies translate from
 * tf_beg to tf_end's translation with tf_beg's rotation: TranslationMotion, math/motion/translation_motion-inl.h:47-84)
 * and request.ccd_solver_type = CCDC_CONSERVATIVE_ADVANCEMENT: BVHConservativeAdvancement<OBBRSS> ->
 * conservativeAdvancementMeshOriented (detail/conservative_advancement_func_matrix-inl.h:149-219, 692-712) with
 * MeshConservativeAdvancementTraversalNodeOBBRSS (detail/traversal/distance/mesh_conservative_advancement_traversal_node.h
 * :163-215, -inl.h:432-713).  Outputs mirror ContinuousCollisionResult: is_collide, time_of_contact (1 when there is
 * no contact), contact_tf1 / contact_tf2 (12 doubles each; the start poses when there is no contact -- the reference
 * leaves them unset).  The other motion types return FCLGPU_ERR_UNSUPPORTED_FUNCTION (see csrc/continuous.cuh), the
 * other solvers as well.  num_max_iterations / toc_err are accepted and ignored, as the reference ignores them on this
 * path (the traversal node's own t_err = 1e-5 ends the advancement).
 * is_collide must not be NULL (it doubles as scratch for the start-configuration verdicts).
 * ------------------------------------------------------------------------------------- */
enum { FCLGPU_CCDM_TRANS = 0, FCLGPU_CCDM_LINEAR = 1, FCLGPU_CCDM_SCREW = 2, FCLGPU_CCDM_SPLINE = 3 };
enum { FCLGPU_CCDC_NAIVE = 0, FCLGPU_CCDC_CONSERVATIVE_ADVANCEMENT = 1, FCLGPU_CCDC_RAY_SHOOTING = 2, FCLGPU_CCDC_POLYNOMIAL_SOLVER = 3 };
typedef struct fclgpu_continuous_request { /* fcl::ContinuousCollisionRequest, continuous_collision_request.h:54-80 */
  int64_t num_max_iterations; /* default 10 (ignored on this path) */
  double toc_err;             /* default 0.0001 (ignored on this path) */
  int32_t ccd_motion_type;    /* FCLGPU_CCDM_* (default CCDM_TRANS) */
  int32_t gjk_solver_type;    /* unused for mesh pairs */
  int32_t ccd_solver_type;    /* FCLGPU_CCDC_* (the reference's default is CCDC_NAIVE: not built) */
} fclgpu_continuous_request;

int fclgpu_continuous_collide_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1_beg,
                                    const double* tf1_end, const double* tf2_beg, const double* tf2_end,
                                    const fclgpu_continuous_request* request, int32_t* is_collide, double* time_of_contact,
                                    double* contact_tf1, double* contact_tf2, int32_t* iterations, uint32_t* n_bv,
                                    uint32_t* n_leaf, void* stream);
int fclgpu_continuous_collide_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1_beg,
                                         const double* tf1_end, const double* tf2_beg, const double* tf2_end,
                                         const fclgpu_continuous_request* request, int32_t* is_collide, double* time_of_contact,
                                         double* contact_tf1, double* contact_tf2, int32_t* iterations);

/* ---------------------------------------------------------------------------------------
 * Batched mesh <-> sphere collide: query i evaluates
 * fcl::collide(m1, tf1[i], Sphere(radius), tf2[i], request, result_i) =
 * BVHShapeCollider<OBBRSS<S>, Sphere<S>> -> orientedBVHShapeCollide
 * (narrowphase/detail/collision_func_matrix-inl.h:378-430, mesh_shape_collision_traversal_node-inl.h:193-262)
 * with the closed-form sphereTriangleIntersect leaf test (sphere_triangle-inl.h:143-244; no GJK on this pair).
 * One contact per intersecting triangle, in the reference's traversal order: b1 = triangle id, b2 = -1
 * (Contact::NONE), pos = contact point, normal = -(centre -> contact) direction, penetration_depth <= 0 as the
 * reference reports it.  Same buffers, capacities and error conventions as fclgpu_collide_batch.
 * ------------------------------------------------------------------------------------- */
int fclgpu_collide_mesh_sphere_batch(const fclgpu_model* m1, double radius, int64_t n, const double* tf1,
                                     const double* tf2, const fclgpu_collision_request* request,
                                     int32_t* num_contacts, fclgpu_contact* contacts, int64_t contact_capacity,
                                     int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf, void* stream);
int fclgpu_collide_mesh_sphere_batch_host(const fclgpu_model* m1, double radius, int64_t n, const double* tf1,
                                          const double* tf2, const fclgpu_collision_request* request,
                                          int32_t* num_contacts, fclgpu_contact* contacts, int64_t contact_capacity,
                                          int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf);

/* ---------------------------------------------------------------------------------------
 * Batched mesh <-> halfspace / plane collide: query i evaluates
 * fcl::collide(m1, tf1[i], Halfspace(normal, d) | Plane(normal, d), tf2[i], request, result_i) =
 * BVHShapeCollider<OBBRSS<S>, Halfspace<S> | Plane<S>> (narrowphase/detail/collision_func_matrix-inl.h:378-430, cells
 * :841-842) with the closed-form leaf tests halfspaceTriangleIntersect (primitive_shape_algorithm/halfspace-inl.h:587-621)
 * and planeTriangleIntersect (plane-inl.h:683-759; no GJK on these pairs, gjk_solver_libccd-inl.h:502-540).
 * (normal, d): the halfspace { x : normal . x <= d } / the plane { normal . x = d } in the shape's frame, normalised like
 * the reference's constructors do (geometry/shape/halfspace-inl.h:144-160).  One contact per intersecting triangle, in the
 * reference's traversal order: b1 = triangle id, b2 = -1 (Contact::NONE); halfspace: pos = deepest vertex moved half the
 * depth back to the boundary, normal = -n', penetration_depth >= 0; plane: pos = middle of the cut segment, normal = -/+ n'.
 * Same buffers, capacities and error conventions as fclgpu_collide_batch.
 * ------------------------------------------------------------------------------------- */
#define FCLGPU_SHAPE_HALFSPACE 17 /* fcl::GEOM_HALFSPACE (geometry/collision_geometry.h:53-54) */
#define FCLGPU_SHAPE_PLANE 16     /* fcl::GEOM_PLANE */
int fclgpu_collide_mesh_plane_batch(const fclgpu_model* m1, int32_t shape, const double* normal3, double d, int64_t n,
                                    const double* tf1, const double* tf2, const fclgpu_collision_request* request,
                                    int32_t* num_contacts, fclgpu_contact* contacts, int64_t contact_capacity,
                                    int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf, void* stream);
int fclgpu_collide_mesh_plane_batch_host(const fclgpu_model* m1, int32_t shape, const double* normal3, double d, int64_t n,
                                         const double* tf1, const double* tf2, const fclgpu_collision_request* request,
                                         int32_t* num_contacts, fclgpu_contact* contacts, int64_t contact_capacity,
                                         int64_t* contact_offsets, uint32_t* n_bv, uint32_t* n_leaf);

/* ---------------------------------------------------------------------------------------
 * Batched distance: query i evaluates fcl::distance(m1, tf1[i], m2, tf2[i], request, result_i)
 * (distance-inl.h:92-246 -> orientedMeshDistance, distance_func_matrix-inl.h:386-403;
 * recursive traversal, qsize = 2, traversal/collision_node.h:67).
 *   min_distance[i]; nearest_p1/p2 (n x 3, world frame, only written when
 *   enable_nearest_points); b1/b2 = closest triangle ids.  Any output may be NULL.
 * ------------------------------------------------------------------------------------- */
int fclgpu_distance_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n,
                          const double* tf1, const double* tf2,
                          const fclgpu_distance_request* request, double* min_distance,
                          double* nearest_p1, double* nearest_p2, int32_t* b1, int32_t* b2,
                          uint32_t* n_bv, uint32_t* n_leaf, void* stream);

int fclgpu_distance_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n,
                               const double* tf1, const double* tf2,
                               const fclgpu_distance_request* request, double* min_distance,
                               double* nearest_p1, double* nearest_p2, int32_t* b1, int32_t* b2,
                               uint32_t* n_bv, uint32_t* n_leaf);

/* ---------------------------------------------------------------------------------------
 * Tolerance verification (EXTENSION; BASELINE cfg5).  The reference advertises "tolerance verification" for
 * meshes (README.md:13-14) but has no API for it; a caller gets it from fcl::distance by comparing the result with
 * the tolerance.  Here the traversal of fclgpu_distance_batch starts from min_distance = cutoff instead of the
 * DistanceResult default DBL_MAX (distance_result-inl.h:52-60), so every node pair whose bound is >= cutoff is pruned
 * from the first round on:
 *   min_distance[i] = the same value fclgpu_distance_batch returns when that is < cutoff (ids and nearest points too),
 *   otherwise cutoff, with b1 = b2 = -1 and unspecified nearest points.
 * "within tolerance tol"  <=>  min_distance[i] <= tol  with cutoff = nextafter(tol, +inf).  cutoff must be > 0.
 * ------------------------------------------------------------------------------------- */
int fclgpu_distance_cutoff_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                 const double* tf2, const fclgpu_distance_request* request, double cutoff,
                                 double* min_distance, double* nearest_p1, double* nearest_p2, int32_t* b1, int32_t* b2,
                                 uint32_t* n_bv, uint32_t* n_leaf, void* stream);
int fclgpu_distance_cutoff_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                      const double* tf2, const fclgpu_distance_request* request, double cutoff,
                                      double* min_distance, double* nearest_p1, double* nearest_p2, int32_t* b1, int32_t* b2,
                                      uint32_t* n_bv, uint32_t* n_leaf);

/* Tolerance VERDICTS (EXTENSION; BASELINE cfg5): within[i] = 1 iff fcl::distance(m1, tf1[i], m2, tf2[i]) <= tolerance,
 * the comparison a caller of the reference would make.  Same traversal as fclgpu_distance_cutoff_batch with
 * cutoff = nextafter(tolerance, +inf), and it ENDS a query at the first triangle pair found within the tolerance
 * (the verdict is decided; the exact minimum is not needed).  witness_distance (optional): the distance of that pair
 * -- an upper bound of the true distance, <= tolerance -- or nextafter(tolerance) when nothing is within.  tolerance >= 0. */
int fclgpu_within_tolerance_batch(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                  const double* tf2, double tolerance, uint8_t* within, double* witness_distance,
                                  uint32_t* n_bv, uint32_t* n_leaf, void* stream);
int fclgpu_within_tolerance_batch_host(const fclgpu_model* m1, const fclgpu_model* m2, int64_t n, const double* tf1,
                                       const double* tf2, double tolerance, uint8_t* within, double* witness_distance,
                                       uint32_t* n_bv, uint32_t* n_leaf);

/* ---------------------------------------------------------------------------------------
 * Batched mesh <-> sphere distance: query i evaluates
 * fcl::distance(m1, tf1[i], Sphere(radius), tf2[i], request, result_i) =
 * BVHShapeDistancer<OBBRSS<S>, Sphere<S>> -> orientedBVHShapeDistance
 * (narrowphase/detail/distance_func_matrix-inl.h:259-277, 322-341;
 * traversal/distance/mesh_shape_distance_traversal_node-inl.h:163-236, 351-413) with the closed-form
 * sphereTriangleDistance leaf (primitive_shape_algorithm/sphere_triangle-inl.h:469-508 on
 * Project::projectTriangle, math/detail/project-inl.h:54-123; no GJK on this pair).
 *   min_distance[i] = min over the triangles of (distance(centre, triangle) - radius);
 *   nearest_p1 = closest point on the mesh in the MESH frame, nearest_p2 = closest point on the sphere in the
 *   SPHERE frame (the reference's postprocess is empty for this node: both stay local), written only when
 *   enable_nearest_points; b1 = closest triangle, b2 = -1 (DistanceResult::NONE).  Any output may be NULL.
 *   Centre within the radius of a triangle: the reference's solver returns false without writing the distance
 *   and its leaf reads it uninitialised; here that case is DEFINED as min_distance = -1 (the value the
 *   distance-only overload writes, sphere_triangle-inl.h:462), NaN points, b1 = a triangle within the radius.
 * ------------------------------------------------------------------------------------- */
int fclgpu_distance_mesh_sphere_batch(const fclgpu_model* m1, double radius, int64_t n, const double* tf1,
                                      const double* tf2, const fclgpu_distance_request* request,
                                      double* min_distance, double* nearest_p1, double* nearest_p2, int32_t* b1,
                                      int32_t* b2, uint32_t* n_bv, uint32_t* n_leaf, void* stream);
int fclgpu_distance_mesh_sphere_batch_host(const fclgpu_model* m1, double radius, int64_t n, const double* tf1,
                                           const double* tf2, const fclgpu_distance_request* request,
                                           double* min_distance, double* nearest_p1, double* nearest_p2, int32_t* b1,
                                           int32_t* b2, uint32_t* n_bv, uint32_t* n_leaf);

/* ---------------------------------------------------------------------------------------
 * Broadphase: N x M culling feeding the batched mesh-mesh kernel -- what
 * NaiveCollisionManager::collide(other_manager, cdata, DefaultCollisionFunction) does
 * (broadphase/broadphase_bruteforce-inl.h:182-205, default_broadphase_callbacks.h:84-103; the dynamic AABB tree manager
 * reports the same set of pairs), for two sets of posed objects that share a table of geometries.
 *   object i of set 1 = (geoms[geom1[i]], tf1[i]), object j of set 2 = (geoms[geom2[j]], tf2[j]);
 *   world AABB of an object = CollisionObject::computeAABB (collision_object-inl.h:118-131): the local AABB translated when
 *   the rotation is the identity, else the cube of half side aabb_radius around tf * aabb_center, with
 *   aabb_local / aabb_center / aabb_radius = BVHModel::computeLocalAABB (BVH_model-inl.h:1080-1100, here over the vertices
 *   the triangles reference; fclgpu_model_local_aabb reads them);
 *   pairs[2k], pairs[2k+1] = (i, j) of the k-th pair whose AABBs overlap (AABB::overlap, AABB-inl.h:98-107), in the
 *   brute-force manager's visiting order (i ascending, then j ascending); *num_pairs = their number (also when it
 *   exceeds pair_capacity: FCLGPU_ERR_CONTACT_OVERFLOW, rerun with a larger list);
 *   num_contacts[k] (optional) = fcl::collide(o1, tf1[i], o2, tf2[j], request, fresh result) for pair k: the pairs are
 *   grouped by geometry pair ON THE DEVICE and run through fclgpu_collide_batch; neither the pair list nor the gathered
 *   poses visit the host in between.  aabb1_out / aabb2_out (optional): the world AABBs (n x 6: min, max).
 * Host pointers; returns when the results are in place.
 * ------------------------------------------------------------------------------------- */
int fclgpu_model_local_aabb(const fclgpu_model* m, double center3[3], double* radius, double min3[3], double max3[3]);
int fclgpu_broadphase_collide_host(int32_t n_geoms, const fclgpu_model* const* geoms, int64_t n1, const int32_t* geom1,
                                   const double* tf1, int64_t n2, const int32_t* geom2, const double* tf2,
                                   const fclgpu_collision_request* request, int64_t pair_capacity, int32_t* pairs,
                                   int64_t* num_pairs, int32_t* num_contacts, double* aabb1_out, double* aabb2_out);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU: the path shards over poses with NO exchange inside the traversal -- BVHs are replicated
 * (upload the models on every GPU), rank r owns the contiguous block fclgpu_shard_range() names, every rank runs the
 * *_batch calls above on its block, and the per-rank result arrays are gathered with NCCL's all-gather over NVLink.
 * One process per GPU.  NCCL is opened at run time (libnccl.so.2); without it these calls return
 * FCLGPU_ERR_UNSUPPORTED_FUNCTION.  Bootstrap: rank 0 calls fclgpu_comm_unique_id and hands the 128 bytes to the other
 * ranks by whatever the application uses to start its processes (MPI, a file, a socket); every rank then calls
 * fclgpu_comm_init.  The collectives take DEVICE pointers and are asynchronous on `stream`: enqueue the gather of
 * batch k on a second stream (ordered by an event) to overlap it with the traversal of batch k + 1.
 * ------------------------------------------------------------------------------------- */
#define FCLGPU_COMM_ID_BYTES 128
typedef struct fclgpu_comm fclgpu_comm;
/* rank `rank` of `world` owns queries [*start, *start + *count) of a batch of n (balanced to within one query) */
void fclgpu_shard_range(int64_t n, int rank, int world, int64_t* start, int64_t* count);
int fclgpu_comm_unique_id(char id[FCLGPU_COMM_ID_BYTES]);
int fclgpu_comm_init(int device, int rank, int world, const char id[FCLGPU_COMM_ID_BYTES], fclgpu_comm** out);
int fclgpu_comm_rank(const fclgpu_comm* comm);
int fclgpu_comm_world(const fclgpu_comm* comm);
/* fixed-size records: recv (world * bytes_per_rank bytes) receives rank r's block at r * bytes_per_rank */
int fclgpu_comm_allgather(fclgpu_comm* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream);
/* ragged blocks (contact lists): rank r contributes bytes_of_rank[r] bytes (HOST array, identical on every rank: gather
 * the totals first with fclgpu_comm_allgather); recv receives the blocks back to back in rank order */
int fclgpu_comm_allgather_ragged(fclgpu_comm* comm, const void* send, void* recv, const int64_t* bytes_of_rank, void* stream);
int fclgpu_comm_destroy(fclgpu_comm* comm);
const char* fclgpu_comm_last_error(void); /* thread-local */

/* ---------------------------------------------------------------------------------------
 * Utilities
 * ------------------------------------------------------------------------------------- */
int fclgpu_abi_version(void);
int fclgpu_device_count(void);
const char* fclgpu_last_error(void); /* thread-local, human readable */
/* Eigen::Transform<double,3,Isometry> (4x4 column-major, 16 doubles) -> 12-double pose */
void fclgpu_pose_from_colmajor4x4(const double* m16, double* pose12);
/* The *_batch entry points are asynchronous; errors detected on the device (contact
 * capacity, traversal stack) are sticky.  This synchronises `stream` and returns and clears
 * that status. */
int fclgpu_sync_status(int device, void* stream);
/* The reference allocates nothing persistent (its traversal nodes live on the stack, SURVEY 8b "Ownership"); this library
 * keeps a per-device workspace that grows with the largest batch seen (contact staging, host-API staging buffers, the
 * distance front's overflow areas, scan temporaries).  fclgpu_device_trim waits for the device to go idle and gives those
 * buffers back; the next call allocates what it needs again.  Models are not touched.  Returns the number of bytes released
 * through *released (may be NULL).  Must not run concurrently with *_batch calls that are still being ENQUEUED on the same
 * device from other threads (it takes the same locks, so it is safe, but it would stall them). */
int fclgpu_device_trim(int device, int64_t* released);
/* Tuning knob: select the traversal kernel variant (0 = default).  See DESIGN.md. */
int fclgpu_set_option(const char* name, int64_t value);
int64_t fclgpu_get_option(const char* name);
/* Kernel launches issued by this library since process start (bench.py's gpu_launches). */
int64_t fclgpu_launch_count(void);
/* Development counters of instrumented builds (-DFCLGPU_DIST_PROF=1: per-phase cycles of the distance kernel); a product
 * build returns zeros. */
int fclgpu_debug_counters(int device, uint64_t* out16, int reset);
/* FP64 pipe / L2 micro-benchmarks used for the roofline denominators:
 * kind 0: unfused DMUL+DADD ops/s, 1: DFMA ops/s (counted as 1 op), 2: L2-resident read GB/s */
int fclgpu_microbench(int device, int kind, double* result);

#ifdef __cplusplus
}
#endif
#endif /* FCLGPU_H_ */
