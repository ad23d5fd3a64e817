// zstd-compressed model images: see zstd_loader.cpp.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "common.hpp"

namespace vpt {

// true when `data` starts with a zstd frame (or skippable frame) magic number
bool is_zstd_frame(const uint8_t* data, size_t len);
// Decodes every frame of `data` (what zstd::Decoder yields when read to the end).  Throws Error(kIoError).
std::vector<uint8_t> zstd_decode_all(const uint8_t* data, size_t len);

}  // namespace vpt
