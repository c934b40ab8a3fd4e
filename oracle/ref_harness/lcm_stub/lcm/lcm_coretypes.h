/*
 * lcm_coretypes.h -- TEST INFRASTRUCTURE: stand-in for the header of the third-party LCM library (lcm-proj/lcm,
 * lcm/lcm_coretypes.h), which the reference's generated message types include but which is absent from /root/reference
 * (no submodule, no pinned version; the reference README only links https://lcm-proj.github.io/).  It restates the
 * published wire primitives the generated code calls: every scalar is written in network byte order (big endian),
 * arrays element by element, sizes in bytes returned.  Parity of these primitives with the real library is UNPINNED;
 * what tests/golden/lcm_traj_f.bin pins is the reference's own generated type (field order, fingerprint constant
 * 0x8fb839bd5c6031ee rotated left by one, lcmtypes/drake/lcmt_trajectory_f.hpp:256-260) and the way
 * LCMHelpers.cuh:245-252 fills it.
 */
#ifndef LCM_CORETYPES_STUB_H
#define LCM_CORETYPES_STUB_H
#include <stdint.h>
#include <string.h>
typedef struct ___lcm_hash_ptr __lcm_hash_ptr;
struct ___lcm_hash_ptr { const __lcm_hash_ptr *parent; int64_t (*v)(void); };
#define LCM_STUB_ARRAY(NAME, T, SZ) \
static inline int __##NAME##_encoded_array_size(const T *p, int elements){ (void)p; return SZ*elements; } \
static inline int __##NAME##_encode_array(void *_buf, int offset, int maxlen, const T *p, int elements){ \
    if (maxlen < SZ*elements){ return -1; } uint8_t *buf = (uint8_t*)_buf + offset; \
    for (int e = 0; e < elements; e++){ uint64_t v = 0; memcpy(&v, &p[e], SZ); for (int b = 0; b < SZ; b++){ buf[e*SZ + b] = (uint8_t)(v >> (8*(SZ-1-b))); } } \
    return SZ*elements; } \
static inline int __##NAME##_decode_array(const void *_buf, int offset, int maxlen, T *p, int elements){ \
    if (maxlen < SZ*elements){ return -1; } const uint8_t *buf = (const uint8_t*)_buf + offset; \
    for (int e = 0; e < elements; e++){ uint64_t v = 0; for (int b = 0; b < SZ; b++){ v = (v << 8) | buf[e*SZ + b]; } memcpy(&p[e], &v, SZ); } \
    return SZ*elements; }
LCM_STUB_ARRAY(int64_t, int64_t, 8)
LCM_STUB_ARRAY(int32_t, int32_t, 4)
LCM_STUB_ARRAY(float, float, 4)
#endif
