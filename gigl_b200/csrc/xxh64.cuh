// xxh64.cuh - XXH64 of one little-endian int32 with seed 42, i.e. Spark's XXH64.hashInt(int, 42L)
// (org.apache.spark:spark-sql 3.1.3, used by `xxhash64(x + _internal_seed + seed)` in
// scala/subgraph_sampler/src/main/scala/libs/task/SamplingStrategy.scala:55).
#pragma once
#include <stdint.h>

namespace gigl {

constexpr uint64_t kP1 = 0x9E3779B185EBCA87ULL;
constexpr uint64_t kP2 = 0xC2B2AE3D27D4EB4FULL;
constexpr uint64_t kP3 = 0x165667B19E3779F9ULL;
constexpr uint64_t kP5 = 0x27D4EB2F165667C5ULL;
constexpr uint64_t kSeed42Init = 42ULL + kP5 + 4ULL;  // seed + PRIME64_5 + len

__host__ __device__ __forceinline__ uint64_t xxh64_int_seed42(int32_t v) {
    uint64_t h = kSeed42Init ^ ((uint64_t)(uint32_t)v * kP1);
    h = ((h << 23) | (h >> 41)) * kP2 + kP3;
    h ^= h >> 33;
    h *= kP2;
    h ^= h >> 29;
    h *= kP3;
    h ^= h >> 32;
    return h;
}

// Order-preserving map signed-int64 order -> unsigned order (Spark compares the hash as a
// signed bigint).  XXH64 restricted to 32-bit inputs is injective (every step is a bijection
// on the zero-extended input) and no int32 input maps to INT64_MAX (checked exhaustively, see
// tests/test_oracle_sampler.py::test_sentinel_unreachable), so ~0ULL is a safe "+infinity".
__host__ __device__ __forceinline__ uint64_t ordered_key(int32_t v) {
    return xxh64_int_seed42(v) ^ 0x8000000000000000ULL;
}
constexpr uint64_t kKeyInf = ~0ULL;

}  // namespace gigl
