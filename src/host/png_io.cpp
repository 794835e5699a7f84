#include "png_io.h"

#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace png_io {
namespace {

enum : unsigned { kOk = 0, kOpen = 1, kSignature = 2, kChunk = 3, kHeader = 4, kUnsupported = 5, kInflate = 6, kSize = 7, kWrite = 8, kEmpty = 9 };

const unsigned char kSig[8] = {137, 80, 78, 71, 13, 10, 26, 10};

uint32_t be32(const unsigned char* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
void put32(std::vector<unsigned char>& v, uint32_t x) {
    v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x);
}

int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool read_file(const std::string& fn, std::vector<unsigned char>& buf) {
    FILE* f = fopen(fn.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
    fclose(f);
    return got == buf.size();
}

void write_chunk(std::vector<unsigned char>& out, const char type[4], const unsigned char* data, size_t n) {
    put32(out, (uint32_t)n);
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    put32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4)));
}

}  // namespace

const char* error_text(unsigned code) {
    switch (code) {
        case kOk: return "ok";
        case kOpen: return "cannot open file";
        case kSignature: return "not a PNG file";
        case kChunk: return "corrupt chunk structure";
        case kHeader: return "bad IHDR";
        case kUnsupported: return "unsupported PNG variant (need 8-bit, non-interlaced)";
        case kInflate: return "zlib inflate failed";
        case kSize: return "decoded size mismatch";
        case kWrite: return "cannot write file";
        case kEmpty: return "empty image";
        default: return "unknown error";
    }
}

unsigned decode(std::vector<unsigned char>& out, unsigned& w, unsigned& h, const std::string& filename) {
    std::vector<unsigned char> file;
    if (!read_file(filename, file)) return kOpen;
    if (file.size() < 8 + 25 || memcmp(file.data(), kSig, 8) != 0) return kSignature;
    size_t pos = 8;
    unsigned bit_depth = 0, color_type = 0, interlace = 0;
    bool have_ihdr = false;
    std::vector<unsigned char> idat, palette, trns;
    while (pos + 12 <= file.size()) {
        uint32_t len = be32(&file[pos]);
        const unsigned char* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) return kChunk;
        const unsigned char* data = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return kHeader;
            w = be32(data); h = be32(data + 4);
            bit_depth = data[8]; color_type = data[9]; interlace = data[12];
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!memcmp(type, "tRNS", 4)) {
            trns.assign(data, data + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || w == 0 || h == 0) return kHeader;
    if (w > 65535u || h > 65535u) return kUnsupported;          // the library limits H to 65535; also bounds the allocation below
    if (bit_depth != 8 || interlace != 0) return kUnsupported;
    unsigned ch;
    switch (color_type) {
        case 0: ch = 1; break;
        case 2: ch = 3; break;
        case 3: ch = 1; break;
        case 4: ch = 2; break;
        case 6: ch = 4; break;
        default: return kUnsupported;
    }
    const size_t stride = (size_t)w * ch;
    std::vector<unsigned char> raw((stride + 1) * h);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK) return kInflate;
    if (raw_len != raw.size()) return kSize;
    // undo the scanline filters in place (PNG spec section 9)
    std::vector<unsigned char> img(stride * h);
    for (unsigned y = 0; y < h; y++) {
        const unsigned char* in = &raw[(stride + 1) * y];
        unsigned char* cur = &img[stride * y];
        const unsigned char* up = y ? &img[stride * (y - 1)] : nullptr;
        const unsigned ft = in[0];
        for (size_t i = 0; i < stride; i++) {
            int a = i >= ch ? cur[i - ch] : 0, b = up ? up[i] : 0, c = (up && i >= ch) ? up[i - ch] : 0;
            int x = in[1 + i];
            switch (ft) {
                case 0: break;
                case 1: x += a; break;
                case 2: x += b; break;
                case 3: x += (a + b) / 2; break;
                case 4: x += paeth(a, b, c); break;
                default: return kChunk;
            }
            cur[i] = (unsigned char)x;
        }
    }
    out.resize((size_t)w * h * 4);
    for (size_t p = 0; p < (size_t)w * h; p++) {
        unsigned char r, g, b, a = 255;
        const unsigned char* s = &img[p * ch];
        switch (color_type) {
            case 0: r = g = b = s[0]; break;
            case 2: r = s[0]; g = s[1]; b = s[2]; break;
            case 3: {
                size_t idx = s[0];
                if (idx * 3 + 2 >= palette.size()) { r = g = b = 0; }
                else { r = palette[idx * 3]; g = palette[idx * 3 + 1]; b = palette[idx * 3 + 2]; }
                if (idx < trns.size()) a = trns[idx];
                break;
            }
            case 4: r = g = b = s[0]; a = s[1]; break;
            default: r = s[0]; g = s[1]; b = s[2]; a = s[3]; break;
        }
        out[p * 4] = r; out[p * 4 + 1] = g; out[p * 4 + 2] = b; out[p * 4 + 3] = a;
    }
    return kOk;
}

unsigned encode(const std::string& filename, const unsigned char* rgba, unsigned w, unsigned h) {
    if (!rgba || w == 0 || h == 0) return kEmpty;
    const size_t stride = (size_t)w * 4;
    std::vector<unsigned char> raw((stride + 1) * h);
    for (unsigned y = 0; y < h; y++) {   // filter type 1 (Sub): cheap and good for flat disparity maps
        unsigned char* o = &raw[(stride + 1) * y];
        const unsigned char* s = rgba + stride * y;
        o[0] = 1;
        for (size_t i = 0; i < stride; i++) o[1 + i] = (unsigned char)(s[i] - (i >= 4 ? s[i - 4] : 0));
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return kInflate;
    std::vector<unsigned char> out(kSig, kSig + 8);
    unsigned char ihdr[13];
    ihdr[0] = w >> 24; ihdr[1] = w >> 16; ihdr[2] = w >> 8; ihdr[3] = w;
    ihdr[4] = h >> 24; ihdr[5] = h >> 16; ihdr[6] = h >> 8; ihdr[7] = h;
    ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    write_chunk(out, "IHDR", ihdr, 13);
    write_chunk(out, "IDAT", z.data(), zlen);
    write_chunk(out, "IEND", nullptr, 0);
    FILE* f = fopen(filename.c_str(), "wb");
    if (!f) return kWrite;
    size_t n = fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    return n == out.size() ? kOk : kWrite;
}

unsigned encode(const std::string& filename, const std::vector<unsigned char>& rgba, unsigned w, unsigned h) {
    if (rgba.size() < (size_t)w * h * 4) return kEmpty;
    return encode(filename, rgba.data(), w, h);
}

}  // namespace png_io
