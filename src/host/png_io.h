// png_io.h -- minimal PNG codec for the host flow, with the call shape of the codec the
// reference uses (lodepng::decode(vector&, w, h, path) / lodepng::encode(path, vector, w, h),
// stereo_matching/main.cpp:184-186,623).  Written from the PNG specification on top of zlib;
// 8-bit non-interlaced images only (every bundled dataset image is 8-bit RGB).
#pragma once
#include <string>
#include <vector>

namespace png_io {

// Decodes to RGBA8 (alpha = 255 when the file has none), like lodepng's default.
// Returns 0 on success, a non-zero error code otherwise (message via error_text()).
unsigned decode(std::vector<unsigned char>& out, unsigned& w, unsigned& h, const std::string& filename);

// Encodes RGBA8 pixels as an 8-bit RGBA PNG.  Returns 0 on success.
unsigned encode(const std::string& filename, const std::vector<unsigned char>& rgba, unsigned w, unsigned h);
unsigned encode(const std::string& filename, const unsigned char* rgba, unsigned w, unsigned h);

const char* error_text(unsigned code);

}  // namespace png_io
