// Host helper shared by the tensor-core kernels: builds a float32 TMA tensor map (128B swizzle, zero OOB fill).
#pragma once
#include <cuda.h>
#include <stdint.h>
namespace sgv {
// atom32 = false: SWIZZLE_128B (16-byte chunks XOR row%8; K-major UMMA operands)
// atom32 = true : SWIZZLE_128B_ATOM_32B (32-byte chunks XOR row%4; the only MN-major layout for 32-bit UMMA operands)
int make_tmap_f32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides, bool atom32 = false);
}
