// tcgen05 (5th-gen tensor core) GEMM core -- placeholder until the 3xTF32 kernel lands.
#pragma once
#include "gast_common.cuh"
namespace gast {
struct TcWeights { float* hi = nullptr; float* lo = nullptr; };
inline bool tc_supported(const GemmP&, int, const TcWeights&) { return false; }
inline int tc_launch(int, cudaStream_t, int, const GemmP&, const TcWeights&) { return (int)cudaErrorNotSupported; }
}  // namespace gast
