// Host side of libtutu_b200: image files for the scene-authoring paths.
//
//   tutu_texture_load   ASCII P3 exactly as the reference's PPMGenerator::loadTexture reads it
//                       (PPMGenerator.hpp:1027-1084: texel = (r / max, g / max, b / max) in fp32), plus what
//                       the reference cannot read: binary P6 and PNG ("reference reads ASCII P3 only", :1050)
//   tutu_write_png      8-bit RGB PNG (the author's TODO next to the binary PPM, README.md:49)
//
// Pure C++ (no CUDA).  PNG uses zlib's inflate / deflate only; chunk parsing, CRC checks, scanline
// unfiltering and the sample conversions are done here.  Not supported: Adam7 interlacing (an error).
#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "tutu_internal.hpp"

using namespace tutu;

namespace {

struct FileCloser {
  void operator()(FILE* f) const {
    if (f) fclose(f);
  }
};

bool read_file(const char* path, std::vector<uint8_t>* out) {
  std::unique_ptr<FILE, FileCloser> f(fopen(path, "rb"));
  if (!f) return false;
  if (fseek(f.get(), 0, SEEK_END) != 0) return false;
  const long n = ftell(f.get());
  if (n < 0 || fseek(f.get(), 0, SEEK_SET) != 0) return false;
  out->resize((size_t)n);
  return n == 0 || fread(out->data(), 1, (size_t)n, f.get()) == (size_t)n;
}

int fail(int code, const std::string& msg) {
  set_error(msg);
  return code;
}

// ---- PPM ------------------------------------------------------------------------------------
// Whitespace-separated tokens; '#' comments (which the reference's `input >> token` parser does not know)
// are accepted between header tokens.
struct PpmCursor {
  const uint8_t* p;
  const uint8_t* end;
  bool token(std::string* out) {
    for (;;) {
      while (p < end && isspace(*p)) ++p;
      if (p < end && *p == '#') {
        while (p < end && *p != '\n') ++p;
        continue;
      }
      break;
    }
    if (p >= end) return false;
    out->clear();
    while (p < end && !isspace(*p)) out->push_back((char)*p++);
    return true;
  }
};

// checkPosInt + std::stoi of the reference (global.hpp:71-85): digits only
bool pos_int(const std::string& s, long* v) {
  if (s.empty() || s.size() > 9) return false;
  long acc = 0;
  for (char c : s) {
    if (c < '0' || c > '9') return false;
    acc = acc * 10 + (c - '0');
  }
  *v = acc;
  return true;
}

int load_ppm(const std::vector<uint8_t>& buf, const char* path, std::vector<float>* rgb, int* w, int* h) {
  PpmCursor c{buf.data(), buf.data() + buf.size()};
  std::string magic, t0, t1, t2;
  if (!c.token(&magic) || !c.token(&t0) || !c.token(&t1) || !c.token(&t2))
    return fail(TUTU_E_IO, std::string("tutu_texture_load: truncated PPM header in ") + path);
  const bool binary = magic == "P6";
  long W, H, M;
  if (!pos_int(t0, &W) || !pos_int(t1, &H) || !pos_int(t2, &M) || W <= 0 || H <= 0 || M <= 0 || M > 65535 ||
      (uint64_t)W * (uint64_t)H > (1ull << 28))
    return fail(TUTU_E_IO, std::string("tutu_texture_load: bad PPM header in ") + path);
  const float max = (float)M;  // float max = std::stoi(b0);
  const size_t n = (size_t)W * H;
  rgb->resize(n * 3);
  if (binary) {
    if (c.p >= c.end) return fail(TUTU_E_IO, std::string("tutu_texture_load: truncated P6 file ") + path);
    ++c.p;  // the single whitespace byte after maxval
    const size_t bps = M > 255 ? 2 : 1;
    if ((size_t)(c.end - c.p) < n * 3 * bps) return fail(TUTU_E_IO, std::string("tutu_texture_load: truncated P6 file ") + path);
    for (size_t i = 0; i < n * 3; ++i) {
      const int v = bps == 1 ? c.p[i] : ((c.p[2 * i] << 8) | c.p[2 * i + 1]);
      (*rgb)[i] = v / max;
    }
  } else {
    for (size_t i = 0; i < n * 3; ++i) {
      long v;
      if (!c.token(&t0) || !pos_int(t0, &v))
        return fail(TUTU_E_IO, std::string("tutu_texture_load: bad or missing sample in P3 file ") + path);
      (*rgb)[i] = (int)v / max;  // Vector3f(r / max, g / max, b / max), PPMGenerator.hpp:1075
    }
  }
  *w = (int)W, *h = (int)H;
  return TUTU_OK;
}

// ---- PNG ------------------------------------------------------------------------------------
const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

int load_png(const std::vector<uint8_t>& buf, const char* path, std::vector<float>* rgb, int* w, int* h) {
  const std::string where = std::string(" in ") + path;
  size_t pos = 8;
  uint32_t W = 0, H = 0;
  int depth = 0, ctype = -1;
  std::vector<uint8_t> idat, palette;
  bool seen_end = false;
  while (!seen_end) {
    if (pos + 12 > buf.size()) return fail(TUTU_E_IO, "tutu_texture_load: truncated PNG" + where);
    const uint32_t len = be32(&buf[pos]);
    const uint8_t* type = &buf[pos + 4];
    if (len > buf.size() || pos + 12 + (size_t)len > buf.size()) return fail(TUTU_E_IO, "tutu_texture_load: truncated PNG chunk" + where);
    const uint8_t* data = &buf[pos + 8];
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, 4 + len);
    if (crc != be32(data + len)) return fail(TUTU_E_IO, "tutu_texture_load: PNG chunk CRC mismatch" + where);
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) return fail(TUTU_E_IO, "tutu_texture_load: bad IHDR" + where);
      W = be32(data), H = be32(data + 4);
      depth = data[8], ctype = data[9];
      if (data[10] != 0 || data[11] != 0) return fail(TUTU_E_IO, "tutu_texture_load: unknown PNG compression / filter method" + where);
      if (data[12] != 0) return fail(TUTU_E_IO, "tutu_texture_load: interlaced PNG files are not supported" + where);
    } else if (!memcmp(type, "PLTE", 4)) {
      palette.assign(data, data + len);
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), data, data + len);
    } else if (!memcmp(type, "IEND", 4)) {
      seen_end = true;
    }
    pos += 12 + (size_t)len;
  }
  int channels;
  switch (ctype) {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: return fail(TUTU_E_IO, "tutu_texture_load: bad PNG colour type" + where);
  }
  const bool depth_ok = ctype == 0   ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                     : (depth == 8 || depth == 16);
  if (!depth_ok || W == 0 || H == 0 || (uint64_t)W * H > (1ull << 28)) return fail(TUTU_E_IO, "tutu_texture_load: bad PNG header" + where);
  if (ctype == 3 && (palette.empty() || palette.size() % 3)) return fail(TUTU_E_IO, "tutu_texture_load: palette PNG without PLTE" + where);
  const size_t bpp_bits = (size_t)channels * depth;
  const size_t stride = ((size_t)W * bpp_bits + 7) / 8;
  const size_t bpp = (bpp_bits + 7) / 8;  // filter distance in bytes (>= 1)
  std::vector<uint8_t> raw((stride + 1) * (size_t)H);
  uLongf raw_len = (uLongf)raw.size();
  if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
    return fail(TUTU_E_IO, "tutu_texture_load: PNG image data does not inflate to the declared size" + where);
  // unfilter in place (PNG spec 9.2)
  std::vector<uint8_t> zero(stride, 0);
  for (uint32_t y = 0; y < H; ++y) {
    uint8_t* line = &raw[(stride + 1) * (size_t)y];
    const int filter = line[0];
    uint8_t* cur = line + 1;
    const uint8_t* up = y ? line - stride : zero.data();  // previous line's data starts at (line - stride - 1) + 1
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
      int v = cur[i];
      switch (filter) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: return fail(TUTU_E_IO, "tutu_texture_load: bad PNG filter type" + where);
      }
      cur[i] = (uint8_t)v;
    }
  }
  rgb->resize((size_t)W * H * 3);
  const float max = (float)((1u << (ctype == 3 ? 8 : depth)) - 1u);  // texel = sample / max like the P3 path
  for (uint32_t y = 0; y < H; ++y) {
    const uint8_t* cur = &raw[(stride + 1) * (size_t)y + 1];
    for (uint32_t x = 0; x < W; ++x) {
      int s[4] = {0, 0, 0, 0};
      if (depth == 16) {
        for (int k = 0; k < channels; ++k) s[k] = (cur[((size_t)x * channels + k) * 2] << 8) | cur[((size_t)x * channels + k) * 2 + 1];
      } else if (depth == 8) {
        for (int k = 0; k < channels; ++k) s[k] = cur[(size_t)x * channels + k];
      } else {  // 1, 2, 4 bits: one channel, most significant bits first
        const size_t bit = (size_t)x * depth;
        s[0] = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
      }
      int r, g, b;
      if (ctype == 3) {
        if ((size_t)s[0] * 3 + 2 >= palette.size())
          return fail(TUTU_E_IO, "tutu_texture_load: palette index out of range" + where);
        r = palette[3 * s[0]], g = palette[3 * s[0] + 1], b = palette[3 * s[0] + 2];
      } else if (channels <= 2) {
        r = g = b = s[0];  // grey (alpha ignored)
      } else {
        r = s[0], g = s[1], b = s[2];  // alpha ignored
      }
      float* o = &(*rgb)[((size_t)y * W + x) * 3];
      o[0] = r / max, o[1] = g / max, o[2] = b / max;
    }
  }
  *w = (int)W, *h = (int)H;
  return TUTU_OK;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24)), v.push_back((uint8_t)(x >> 16)), v.push_back((uint8_t)(x >> 8)), v.push_back((uint8_t)x);
}
void put_chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t len) {
  put_be32(out, (uint32_t)len);
  const size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  if (len) out.insert(out.end(), data, data + len);
  put_be32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), &out[start], (uInt)(4 + len)));
}

}  // namespace

// normal_map != 0 applies the recovery the reference applies to a freshly loaded `bump` map
// (PPMGenerator.hpp:714-720): c = c * 2.f, then each component - 1.f.
extern "C" int tutu_texture_load(const char* path, int normal_map, float** rgb_out, int32_t* width, int32_t* height) {
  if (!path || !rgb_out || !width || !height) return fail(TUTU_E_INVALID, "tutu_texture_load: null argument");
  *rgb_out = nullptr;
  try {
    std::vector<uint8_t> buf;
    if (!read_file(path, &buf)) return fail(TUTU_E_IO, std::string("tutu_texture_load: cannot read ") + path);
    std::vector<float> rgb;
    int w = 0, h = 0, rc;
    if (buf.size() >= 8 && !memcmp(buf.data(), kPngSig, 8))
      rc = load_png(buf, path, &rgb, &w, &h);
    else if (buf.size() >= 2 && buf[0] == 'P' && (buf[1] == '3' || buf[1] == '6'))
      rc = load_ppm(buf, path, &rgb, &w, &h);
    else
      return fail(TUTU_E_IO, std::string("tutu_texture_load: ") + path + " is neither a P3 / P6 PPM nor a PNG file");
    if (rc != TUTU_OK) return rc;
    if (normal_map)
      for (float& c : rgb) c = c * 2.f - 1.f;
    float* out = static_cast<float*>(malloc(rgb.size() * sizeof(float) + 1));
    if (!out) return fail(TUTU_E_NOMEM, "tutu_texture_load: out of memory");
    memcpy(out, rgb.data(), rgb.size() * sizeof(float));
    *rgb_out = out;
    *width = w, *height = h;
    return TUTU_OK;
  } catch (const std::bad_alloc&) {
    return fail(TUTU_E_NOMEM, "tutu_texture_load: out of memory");
  } catch (const std::exception& e) {
    return fail(TUTU_E_INVALID, std::string("tutu_texture_load: ") + e.what());
  }
}

extern "C" void tutu_texture_free(float* rgb) { free(rgb); }

extern "C" int tutu_write_png(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb8) {
  if (!path || (!rgb8 && (size_t)width * height != 0) || width == 0 || height == 0)
    return fail(TUTU_E_INVALID, "tutu_write_png: bad argument");
  try {
    const size_t stride = (size_t)width * 3;
    std::vector<uint8_t> raw((stride + 1) * (size_t)height);
    for (uint32_t y = 0; y < height; ++y) {
      raw[(stride + 1) * (size_t)y] = 0;  // filter type None
      memcpy(&raw[(stride + 1) * (size_t)y + 1], rgb8 + stride * y, stride);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return fail(TUTU_E_IO, "tutu_write_png: deflate failed");
    std::vector<uint8_t> out(kPngSig, kPngSig + 8);
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width), put_be32(ihdr, height);
    const uint8_t tail[5] = {8, 2, 0, 0, 0};  // 8 bits, RGB, deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), tail, tail + 5);
    put_chunk(out, "IHDR", ihdr.data(), ihdr.size());
    put_chunk(out, "IDAT", z.data(), zlen);
    put_chunk(out, "IEND", nullptr, 0);
    std::unique_ptr<FILE, FileCloser> f(fopen(path, "wb"));
    if (!f) return fail(TUTU_E_IO, std::string("tutu_write_png: cannot open ") + path);
    if (fwrite(out.data(), 1, out.size(), f.get()) != out.size()) return fail(TUTU_E_IO, std::string("tutu_write_png: short write to ") + path);
    return TUTU_OK;
  } catch (const std::bad_alloc&) {
    return fail(TUTU_E_NOMEM, "tutu_write_png: out of memory");
  } catch (const std::exception& e) {
    return fail(TUTU_E_INVALID, std::string("tutu_write_png: ") + e.what());
  }
}
