// TMA tensor-map construction shared by the CUDA translation units (implemented in gemm/gemm.cu).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace tdp {

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row stride `ld`
// elements, box {box_inner, box_outer}, 128-byte swizzle (box_inner * 2 bytes <= 128).  Cached.
bool make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                  uint32_t box_inner, uint32_t box_outer);

}  // namespace tdp
