/* oracle/ufo_oracle.h -- TEST INFRASTRUCTURE ONLY (see ufo_oracle.c).
 *
 * C interface of the CPU restatement of UFOMap's point-cloud integration path.
 * The entry points mirror oracle/ref_harness.cpp (ufo_ref_*) one to one so the
 * Python test wrapper (tests/oracle_lib.py) can drive either library.
 */
#ifndef UFO_ORACLE_H
#define UFO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

void* ufo_oracle_create(double resolution, unsigned depth_levels, int automatic_pruning,
                        double occupied_thres, double free_thres, double prob_hit,
                        double prob_miss, double clamp_min, double clamp_max, int color);
void ufo_oracle_destroy(void* h);

/* xyz: n*3 doubles, rgb: n*3 bytes or NULL.  Returns wall seconds, <0 on error. */
double ufo_oracle_insert(void* h, const double* origin, const double* xyz, const uint8_t* rgb,
                         size_t n, double max_range, unsigned depth, int simple,
                         unsigned early_stopping, int discrete, int async_unused);

/* leaves!=0: nodes without children; leaves==0: nodes with children. */
size_t ufo_oracle_walk(void* h, int leaves);
void ufo_oracle_walk_fetch(void* h, uint64_t* codes, uint32_t* depths, float* occ, uint8_t* rgb,
                           uint8_t* flags);

int ufo_oracle_node(void* h, uint64_t code, unsigned depth, float* occ, uint8_t* rgb,
                    uint8_t* flags, unsigned* found_depth);

size_t ufo_oracle_compute_ray(void* h, const double* origin, const double* end, double max_range,
                              unsigned depth, uint64_t* codes, size_t cap);
/* castRay as intended (the reference's does not compile, see ufo_oracle.c); 1 = hit */
int ufo_oracle_cast_ray(void* h, const double* origin, const double* direction, int ignore_unknown,
                        double max_range, unsigned depth, uint64_t* code);
size_t ufo_oracle_free_set(void* h, const double* origin, const double* ends, size_t n,
                           unsigned depth, int simple, unsigned early_stopping, uint64_t* codes,
                           size_t cap);

void ufo_oracle_to_key(void* h, const double* xyz, unsigned depth, uint32_t* key);
uint64_t ufo_oracle_to_code(void* h, const double* xyz, unsigned depth);
uint64_t ufo_oracle_key_to_code(const uint32_t* key, unsigned depth);
void ufo_oracle_code_to_key(uint64_t code, unsigned depth, uint32_t* key);
void ufo_oracle_key_to_coord(void* h, const uint32_t* key, unsigned depth, double* xyz);
int ufo_oracle_move_line_inside(void* h, double* a, double* b);
void ufo_oracle_change_bbox(void* h, double* mn, double* mx);
void ufo_oracle_reset_change_bbox(void* h);
void ufo_oracle_sensor_model(void* h, double* out6);
size_t ufo_oracle_memory_usage(void* h);

/* Work counters of the last insert: [0]=rays cast, [1]=DDA visits V,
 * [2]=unique free codes U, [3]=unique hit codes U_h. */
void ufo_oracle_last_counters(void* h, uint64_t* out4);

/* cloud frame: pose7 = tx ty tz qw qx qy qz (Pose6::transform, Pose6(x,y,z,roll,pitch,yaw)) */
void ufo_oracle_transform(const double* pose7, const double* xyz, size_t n, double* out);
void ufo_oracle_pose_from_rpy(double x, double y, double z, double roll, double pitch, double yaw,
                              double* pose7);

/* Octree::write(ostream, compress = false): complete file image; returns its size, copies it
 * when it fits into cap */
size_t ufo_oracle_write(void* h, uint8_t* buf, size_t cap);

/* test helper: collapse every collapsible node (canonical minimal tree of the value field) */
void ufo_oracle_canonicalize(void* h);

/* Octree::writeData(stream, AABB or none, compress = false, min_depth): the node stream only */
size_t ufo_oracle_write_data(void* h, const double* box6, unsigned min_depth, uint8_t* buf, size_t cap);

/* setValueVolume(AABB(min, max), occupancy probability, min_depth) */
void ufo_oracle_set_value_volume(void* h, const double* box6, double occupancy, unsigned min_depth);

/* Octree::readData(stream, AABB or none): merge a node stream into the map; 1 = ok */
int ufo_oracle_read_data(void* h, const double* box6, const uint8_t* buf, size_t size);

#ifdef __cplusplus
}
#endif
#endif
