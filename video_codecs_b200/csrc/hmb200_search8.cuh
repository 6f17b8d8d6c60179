// Tiled byte-SIMD full search for 8-bit planes (placeholder: schedule is empty, the generic kernel runs).
#pragma once
#include <string>
#include <vector>
#include <cuda.h>
#include "hmb200_device.cuh"
#include "hmb200_generic.cuh"

namespace hmb200 {
struct Search8Maps { CUtensorMap map[1]; };
struct Search8Schedule { int n_groups = 0; };
inline bool search8_make_maps(EncodeTiledFn, const DevPlane&, Search8Maps*, std::string*) { return true; }
inline int  search8_configure(std::string*) { return HMB200_OK; }
inline bool search8_build_schedule(const std::vector<SearchTask>&, int, cudaStream_t, Search8Schedule*, std::string*) { return true; }
inline void search8_free_schedule(Search8Schedule*) {}
inline int  search8_launch(const Search8Schedule&, const SearchTask*, hmb200_pu_result*, const DevPlane&, const DevPlane&,
                           const Search8Maps&, int, cudaStream_t) { return 0; }
}
