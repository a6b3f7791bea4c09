#pragma once
#include <cstdint>
namespace tfh {
uint8_t* png_load(const char* path, int* W, int* H);  // malloc'ed RGB8, nullptr on failure
int png_save(const char* path, const uint8_t* rgb, int W, int H);
}  // namespace tfh
