// Host front-end: reads the reference's scene.json (+ buffers) and method JSON and produces the
// plain-array AkrSceneDesc.  Re-implements, for the formats the hot path needs, what the Rust
// host does in:
//   crates/akari_scenegraph/src/scene.rs:598-668   (MmapScene::open, buffer views)
//   crates/akari_render/src/load.rs:129-237,457-535 (transforms, camera, instances, mesh slices)
//   crates/akari_render/src/svm/compiler.rs:25-347  (shader graph -> bytecode + constant blob)
//   crates/akari_integrator/src/lib.rs:57-109        (RenderTask JSON)
// Written from the behaviour of those files; no code is shared with them.
#include "../../../include/akari_b200_host.h"
#include "json.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <cstring>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace {

thread_local std::string g_last_error;

int set_error(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}

std::string read_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open '" + path + "'");
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

bool file_exists(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    return static_cast<bool>(f);
}

std::string dirname_of(const std::string &p) {
    size_t i = p.find_last_of('/');
    if (i == std::string::npos) return ".";
    if (i == 0) return "/";
    return p.substr(0, i);
}

std::string basename_any(const std::string &p) {
    size_t i = p.find_last_of("/\\");
    return i == std::string::npos ? p : p.substr(i + 1);
}

// ---- column-major 4x4, following glam::Mat4's scalar formulas --------------------------------
struct Mat4 {
    float c[4][4];  // c[col][row]
    static Mat4 identity() {
        Mat4 m{};
        for (int i = 0; i < 4; ++i) m.c[i][i] = 1.0f;
        return m;
    }
    static Mat4 from_scale(float x, float y, float z) {
        Mat4 m = identity();
        m.c[0][0] = x;
        m.c[1][1] = y;
        m.c[2][2] = z;
        return m;
    }
    static Mat4 from_translation(float x, float y, float z) {
        Mat4 m = identity();
        m.c[3][0] = x;
        m.c[3][1] = y;
        m.c[3][2] = z;
        return m;
    }
    // glam::Mat4::from_axis_angle
    static Mat4 from_axis_angle(float ax, float ay, float az, float angle) {
        float s = std::sin(angle), co = std::cos(angle);
        float axs = ax * s, ays = ay * s, azs = az * s;
        float axq = ax * ax, ayq = ay * ay, azq = az * az;
        float omc = 1.0f - co;
        float xyomc = ax * ay * omc, xzomc = ax * az * omc, yzomc = ay * az * omc;
        Mat4 m{};
        m.c[0][0] = axq * omc + co;
        m.c[0][1] = xyomc + azs;
        m.c[0][2] = xzomc - ays;
        m.c[1][0] = xyomc - azs;
        m.c[1][1] = ayq * omc + co;
        m.c[1][2] = yzomc + axs;
        m.c[2][0] = xzomc + ays;
        m.c[2][1] = yzomc - axs;
        m.c[2][2] = azq * omc + co;
        m.c[3][3] = 1.0f;
        return m;
    }
    // glam mul_mat4: result.col_j = a.col0*b[j].x + a.col1*b[j].y + a.col2*b[j].z + a.col3*b[j].w
    Mat4 operator*(const Mat4 &b) const {
        Mat4 r{};
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) {
                float v = c[0][i] * b.c[j][0];
                v = v + c[1][i] * b.c[j][1];
                v = v + c[2][i] * b.c[j][2];
                v = v + c[3][i] * b.c[j][3];
                r.c[j][i] = v;
            }
        return r;
    }
    Mat4 transpose() const {
        Mat4 r{};
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) r.c[j][i] = c[i][j];
        return r;
    }
    void store(float out[16]) const { std::memcpy(out, c, sizeof(float) * 16); }
};

constexpr float kPi = 3.14159265358979323846f;

// load.rs:129-171
Mat4 load_transform(const akr::json::Value &t, bool is_camera) {
    const std::string &type = t.at("type").string();
    const akr::json::Value &data = t.at("data");
    if (type == "trs") {
        const std::string &cs = data.at("coordinate_system").string();
        float tr[3], r[3], s[3];
        for (int i = 0; i < 3; ++i) {
            tr[i] = data.at("translation").at(i).f32();
            r[i] = data.at("rotation").at(i).f32();
            s[i] = data.at("scale").at(i).f32();
        }
        Mat4 m = Mat4::identity();
        if (!is_camera) m = Mat4::from_scale(s[0], s[1], s[2]) * m;
        if (cs == "Akari") {
            m = Mat4::from_axis_angle(0, 0, 1, r[2]) * m;
            m = Mat4::from_axis_angle(1, 0, 0, r[0]) * m;
            m = Mat4::from_axis_angle(0, 1, 0, r[1]) * m;
            m = Mat4::from_translation(tr[0], tr[1], tr[2]) * m;
        } else if (cs == "Blender") {
            if (is_camera) m = Mat4::from_axis_angle(1, 0, 0, -kPi / 2.0f) * m;
            m = Mat4::from_axis_angle(1, 0, 0, r[0]) * m;
            m = Mat4::from_axis_angle(0, 0, 1, -r[1]) * m;
            m = Mat4::from_axis_angle(0, 1, 0, r[2]) * m;
            m = Mat4::from_translation(tr[0], tr[2], -tr[1]) * m;
        } else {
            throw std::runtime_error("unknown coordinate_system '" + cs + "'");
        }
        return m;
    }
    if (type == "matrix") {
        // Mat4::from_cols_array_2d(m).transpose(): the JSON rows are matrix rows.
        Mat4 m{};
        for (int row = 0; row < 4; ++row)
            for (int col = 0; col < 4; ++col) m.c[col][row] = data.at(row).at(col).f32();
        return m;
    }
    throw std::runtime_error("unknown transform type '" + type + "'");
}

// ---- base64 (Buffer::EmbeddedBase64) -----------------------------------------------------------
std::vector<uint8_t> base64_decode(const std::string &in) {
    static int8_t lut[256];
    static bool init = false;
    if (!init) {
        std::memset(lut, -1, sizeof(lut));
        const char *abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; ++i) lut[static_cast<uint8_t>(abc[i])] = static_cast<int8_t>(i);
        init = true;
    }
    std::vector<uint8_t> out;
    uint32_t acc = 0;
    int bits = 0;
    for (unsigned char ch : in) {
        if (ch == '=') break;
        int8_t v = lut[ch];
        if (v < 0) continue;
        acc = (acc << 6) | static_cast<uint32_t>(v);
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back(static_cast<uint8_t>((acc >> bits) & 0xFF));
        }
    }
    return out;
}

// ---- SVM compiler (svm/compiler.rs) --------------------------------------------------------------
struct CompiledShader {
    std::vector<AkrSvmNode> nodes;
    std::vector<uint8_t> data;
};

bool same_bytecode(const std::vector<AkrSvmNode> &a, const std::vector<AkrSvmNode> &b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) {
        if (a[i].op != b[i].op || a[i].n_args != b[i].n_args) return false;
        for (uint32_t k = 0; k < a[i].n_args; ++k)
            if (a[i].a[k] != b[i].a[k]) return false;
    }
    return true;
}

// (image, sampler) -> texture index: what SceneLoader::preload collects into the bindless heap (load.rs:494-529,611-646)
using ImageResolver = std::function<uint32_t(const akr::json::Value &image)>;

// ---- image decoding (load.rs:550-610) ---------------------------------------------------------------------------------
// raw float: width * height * channels f32, missing channels filled with 0 (alpha: 1), NOT flipped (load.rs:556-588);
// png / tiff: decoded, flipped vertically, converted to RGBA8 (load.rs:590-603); exr: decoded, flipped, RGBA32F (decode_exr).
// jpeg: decode_jpeg below (lossy format: texel-exactness against jpeg-decoder 0.3 is ASSUMED).  dds needs a decoder this host
// does not carry: AKR_ERR_UNSUPPORTED; tga / bmp are not loadable by the reference either (load.rs:597 `unreachable!()`).
// header sanity before any allocation: a corrupt size field must not turn into a multi-gigabyte request
// (no codec here packs more than ~2 K texels into a byte — deflate tops out at 1032 : 1, a DC-only JPEG block spends 2 bits on
// 64 samples — so the file length bounds the plausible size)
void check_image_size(uint32_t width, uint32_t height, size_t file_bytes, const char *what) {
    const uint64_t texels = (uint64_t)width * height;
    if (width == 0 || height == 0 || width > 65535u || height > 65535u || texels > (1ull << 28) || texels > (uint64_t)file_bytes * 2048u + 4096u)
        throw std::runtime_error(std::string(what) + ": image size " + std::to_string(width) + " x " + std::to_string(height) + " is out of range for a file of " +
                                 std::to_string(file_bytes) + " bytes");
}
uint8_t paeth(uint8_t a, uint8_t b, uint8_t c) {
    int p = (int)a + (int)b - (int)c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void decode_png(const uint8_t *data, size_t len, uint32_t &width, uint32_t &height, std::vector<uint8_t> &rgba8) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (len < 8 || std::memcmp(data, sig, 8) != 0) throw std::runtime_error("png: bad signature");
    size_t pos = 8;
    uint32_t bit_depth = 0, color_type = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    bool have_ihdr = false;
    while (pos + 12 <= len) {
        const uint32_t clen = be32(data + pos);
        const std::string type(reinterpret_cast<const char *>(data + pos + 4), 4);
        if (pos + 12 + clen > len) throw std::runtime_error("png: truncated chunk");
        const uint8_t *body = data + pos + 8;
        if (type == "IHDR") {
            width = be32(body);
            height = be32(body + 4);
            bit_depth = body[8];
            color_type = body[9];
            interlace = body[12];
            have_ihdr = true;
        } else if (type == "PLTE") {
            palette.assign(body, body + clen);
        } else if (type == "tRNS") {
            trns.assign(body, body + clen);
        } else if (type == "IDAT") {
            idat.insert(idat.end(), body, body + clen);
        } else if (type == "IEND") {
            break;
        }
        pos += 12 + clen;
    }
    if (!have_ihdr || width == 0 || height == 0) throw std::runtime_error("png: no IHDR");
    check_image_size(width, height, len, "png");
    if (interlace != 0) throw std::runtime_error("png: interlaced images are not supported");
    if (bit_depth != 8 && bit_depth != 16) throw std::runtime_error("png: only 8 / 16 bits per channel are supported");
    uint32_t ch;
    switch (color_type) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 3: ch = 1; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: throw std::runtime_error("png: bad colour type");
    }
    if (color_type == 3 && bit_depth != 8) throw std::runtime_error("png: palette images must be 8 bit");
    const size_t bpp = (size_t)ch * bit_depth / 8, stride = (size_t)width * bpp;
    std::vector<uint8_t> raw((stride + 1) * height);
    uLongf out_len = static_cast<uLongf>(raw.size());
    if (uncompress(raw.data(), &out_len, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || out_len != raw.size())
        throw std::runtime_error("png: inflate failed");
    std::vector<uint8_t> img(stride * height);
    for (uint32_t y = 0; y < height; ++y) {
        const uint8_t filter = raw[y * (stride + 1)];
        const uint8_t *src = raw.data() + y * (stride + 1) + 1;
        uint8_t *dst = img.data() + y * stride;
        const uint8_t *up = y ? dst - stride : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const uint8_t a = x >= bpp ? dst[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            uint8_t v = src[x];
            switch (filter) {
            case 0: break;
            case 1: v = static_cast<uint8_t>(v + a); break;
            case 2: v = static_cast<uint8_t>(v + b); break;
            case 3: v = static_cast<uint8_t>(v + ((a + b) >> 1)); break;
            case 4: v = static_cast<uint8_t>(v + paeth(a, b, c)); break;
            default: throw std::runtime_error("png: bad filter type");
            }
            dst[x] = v;
        }
    }
    rgba8.resize((size_t)width * height * 4);
    for (uint32_t y = 0; y < height; ++y) {
        const uint32_t fy = height - 1 - y;  // DynamicImage::flipv (load.rs:590)
        for (uint32_t x = 0; x < width; ++x) {
            const uint8_t *px = img.data() + y * stride + x * bpp;
            // 16 -> 8 bit as DynamicImage::to_rgba8 does it: round(v * 255 / 65535) = (v + 128) / 257
            // (image 0.24.7, color.rs `impl FromPrimitive<u16> for u8`; ASSUMED from the pinned crate version, Cargo.lock:1034)
            auto sample = [&](uint32_t c) -> uint8_t {
                return bit_depth == 8 ? px[c] : static_cast<uint8_t>(((((uint32_t)px[2 * c] << 8) | px[2 * c + 1]) + 128u) / 257u);
            };
            uint8_t r, g, b, a = 255;
            if (color_type == 3) {
                const uint32_t i = px[0];
                if (3 * i + 2 >= palette.size()) throw std::runtime_error("png: palette index out of range");
                r = palette[3 * i];
                g = palette[3 * i + 1];
                b = palette[3 * i + 2];
                if (i < trns.size()) a = trns[i];
            } else if (ch <= 2) {
                r = g = b = sample(0);
                if (ch == 2) a = sample(1);
            } else {
                r = sample(0);
                g = sample(1);
                b = sample(2);
                if (ch == 4) a = sample(3);
            }
            uint8_t *o = rgba8.data() + ((size_t)fy * width + x) * 4;
            o[0] = r; o[1] = g; o[2] = b; o[3] = a;
        }
    }
}
// ---- JPEG (jpeg-decoder 0.3 behind image::ImageFormat::Jpeg -> flipv -> to_rgba8, load.rs:590-603) ---------------------
// Sequential and progressive Huffman JPEG (any scan structure), 8 bit, 1 (gray) or 3 (YCbCr) components, sampling factors
// 1 or 2, restart intervals.  JPEG is lossy and the decoded bytes depend on the decoder's arithmetic: this one follows the published
// structure of jpeg-decoder 0.3 — the integer IDCT it took from stb_image (12-bit constants, 512 / 65536 + 128 << 17
// rounding), its triangle-filter ("fancy") chroma upsampling for h2v1 / h1v2 / h2v2, and the 20-bit fixed-point BT.601
// conversion — from memory of that crate, which is not vendored here: texel-exact agreement with the reference is
// ASSUMED, not verified (the test bounds the distance to libjpeg-turbo instead).  Arithmetic-coded, lossless, 12-bit and
// CMYK files are rejected.
struct JpegHuff {
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17] = {0}, maxcode[18] = {0}, valptr[17] = {0};
    bool present = false;
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};
struct JpegBits {
    const uint8_t *p, *end;
    uint32_t acc = 0;
    int n = 0;
    bool marker = false;
    int bit() {
        if (n == 0) {
            uint8_t b = 0;
            if (!marker && p < end) {
                b = *p++;
                if (b == 0xff) {
                    if (p < end && *p == 0x00) ++p;  // stuffed zero
                    else {                            // a marker: feed zeros from here on
                        marker = true;
                        --p;
                        b = 0;
                    }
                }
            }
            acc = b;
            n = 8;
        }
        --n;
        return (acc >> n) & 1;
    }
    int receive(int s) {
        int v = 0;
        for (int i = 0; i < s; ++i) v = (v << 1) | bit();
        return v;
    }
    int decode(const JpegHuff &h) {
        int code = 0;
        for (int l = 1; l <= 16; ++l) {
            code = (code << 1) | bit();
            if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
        }
        throw std::runtime_error("jpeg: bad Huffman code");
    }
    void reset() {
        n = 0;
        acc = 0;
        marker = false;
    }
};
inline int jpeg_extend(int v, int s) { return s && v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
inline uint8_t jpeg_clamp(int64_t x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }
// stb_image's stbi__idct_block (the IDCT jpeg-decoder's idct.rs is a port of): coefficients already dequantised
// (64-bit intermediates: the same values as the 32-bit original on every real image, and no signed overflow on corrupt
// files whose 16-bit quantisation entries blow the coefficients up — found by fuzzing under UBSan)
void jpeg_idct(const int *d, uint8_t *out, size_t stride) {
    auto f2f = [](double x) { return (int64_t)(x * 4096 + 0.5); };
    int64_t val[64];
#define AKR_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                                            \
    int64_t t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                                                \
    p2 = s2; p3 = s6;                                                                                          \
    p1 = (p2 + p3) * f2f(0.5411961);                                                                           \
    t2 = p1 + p3 * f2f(-1.847759065);                                                                          \
    t3 = p1 + p2 * f2f(0.765366865);                                                                           \
    p2 = s0; p3 = s4;                                                                                          \
    t0 = (p2 + p3) * 4096; t1 = (p2 - p3) * 4096;                                                              \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                                                    \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                                                        \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                                                    \
    p5 = (p3 + p4) * f2f(1.175875602);                                                                         \
    t0 = t0 * f2f(0.298631336); t1 = t1 * f2f(2.053119869); t2 = t2 * f2f(3.072711026); t3 = t3 * f2f(1.501321110); \
    p1 = p5 + p1 * f2f(-0.899976223); p2 = p5 + p2 * f2f(-2.562915447);                                        \
    p3 = p3 * f2f(-1.961570560); p4 = p4 * f2f(-0.390180644);                                                  \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;
    for (int i = 0; i < 8; ++i) {  // columns
        const int *c = d + i;
        int64_t *v = val + i;
        if (c[8] == 0 && c[16] == 0 && c[24] == 0 && c[32] == 0 && c[40] == 0 && c[48] == 0 && c[56] == 0) {
            const int64_t dc = (int64_t)c[0] * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
        } else {
            AKR_IDCT_1D(c[0], c[8], c[16], c[24], c[32], c[40], c[48], c[56])
            x0 += 512; x1 += 512; x2 += 512; x3 += 512;
            v[0] = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10;
            v[8] = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10;
            v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
        }
    }
    for (int i = 0; i < 8; ++i) {  // rows
        const int64_t *v = val + i * 8;
        uint8_t *o = out + (size_t)i * stride;
        AKR_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
        o[0] = jpeg_clamp((x0 + t3) >> 17); o[7] = jpeg_clamp((x0 - t3) >> 17);
        o[1] = jpeg_clamp((x1 + t2) >> 17); o[6] = jpeg_clamp((x1 - t2) >> 17);
        o[2] = jpeg_clamp((x2 + t1) >> 17); o[5] = jpeg_clamp((x2 - t1) >> 17);
        o[3] = jpeg_clamp((x3 + t0) >> 17); o[4] = jpeg_clamp((x3 - t0) >> 17);
    }
#undef AKR_IDCT_1D
}
void decode_jpeg(const uint8_t *data, size_t len, uint32_t &width, uint32_t &height, std::vector<uint8_t> &rgba8) {
    static const uint8_t zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                       41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                       30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    if (len < 4 || data[0] != 0xff || data[1] != 0xd8) throw std::runtime_error("jpeg: missing SOI");
    struct Comp {
        int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
        int64_t pred = 0;           // DC predictor (64 bit: cannot overflow on a corrupt stream)
        uint32_t bw = 0, bh = 0;    // blocks per line / column, padded to whole MCUs (the layout of `coef` and `plane`)
        uint32_t rbw = 0, rbh = 0;  // blocks that carry image samples (what a non-interleaved scan visits)
        std::vector<int16_t> coef;  // quantised coefficients, natural order, 64 per block: scans accumulate here
        std::vector<uint8_t> plane;
    };
    std::vector<Comp> comps;
    uint16_t qt[4][64] = {{0}};
    bool have_qt[4] = {false, false, false, false};
    JpegHuff dc[4], ac[4];
    uint32_t restart = 0;
    int hmax = 1, vmax = 1;
    bool have_frame = false, progressive = false, decoded = false;
    int adobe_transform = -1;  // no APP14 marker: three components are YCbCr (JFIF)
    size_t pos = 2;
    auto be16 = [&](size_t p) -> uint32_t {
        if (p + 2 > len) throw std::runtime_error("jpeg: truncated file");
        return ((uint32_t)data[p] << 8) | data[p + 1];
    };
    while (pos + 4 <= len) {
        if (data[pos] != 0xff) {  // entropy-coded bytes left over by a scan that ended on a marker: skip to the next marker
            ++pos;
            continue;
        }
        const uint8_t m = data[pos + 1];
        pos += 2;
        if (m == 0xff) {  // fill byte
            --pos;
            continue;
        }
        if (m == 0xd9) break;
        if (m == 0x00 || m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;
        const uint32_t seglen = be16(pos);
        if (seglen < 2 || pos + seglen > len) throw std::runtime_error("jpeg: bad segment length");
        const uint8_t *seg = data + pos + 2;
        const size_t n = seglen - 2;
        if (m == 0xdb) {  // DQT
            size_t i = 0;
            while (i < n) {
                const int pq = seg[i] >> 4, tq = seg[i] & 15;
                ++i;
                if (tq > 3 || i + (pq ? 128 : 64) > n) throw std::runtime_error("jpeg: bad DQT");
                for (int k = 0; k < 64; ++k) {
                    qt[tq][zigzag[k]] = pq ? (uint16_t)((seg[i] << 8) | seg[i + 1]) : seg[i];
                    i += pq ? 2 : 1;
                }
                have_qt[tq] = true;
            }
        } else if (m == 0xc4) {  // DHT
            size_t i = 0;
            while (i < n) {
                const int tc = seg[i] >> 4, th = seg[i] & 15;
                ++i;
                if (tc > 1 || th > 3 || i + 16 > n) throw std::runtime_error("jpeg: bad DHT");
                JpegHuff &h = tc ? ac[th] : dc[th];
                int total = 0;
                for (int l = 1; l <= 16; ++l) {
                    h.bits[l] = seg[i + l - 1];
                    total += h.bits[l];
                }
                i += 16;
                if (total > 256 || i + total > n) throw std::runtime_error("jpeg: bad DHT");
                std::memcpy(h.vals, seg + i, total);
                i += total;
                h.build();
            }
        } else if (m == 0xc0 || m == 0xc1 || m == 0xc2) {  // SOF0 / SOF1 (sequential), SOF2 (progressive)
            if (have_frame) throw std::runtime_error("jpeg: more than one frame");
            if (n < 6 || seg[0] != 8) throw std::runtime_error("jpeg: only 8-bit samples are supported");
            progressive = m == 0xc2;
            height = (seg[1] << 8) | seg[2];
            width = (seg[3] << 8) | seg[4];
            const int nc = seg[5];
            if ((nc != 1 && nc != 3) || n < (size_t)(6 + 3 * nc) || !width || !height) throw std::runtime_error("jpeg: unsupported component count / size");
            check_image_size(width, height, len, "jpeg");
            comps.resize(nc);
            for (int c = 0; c < nc; ++c) {
                comps[c].id = seg[6 + 3 * c];
                comps[c].h = seg[7 + 3 * c] >> 4;
                comps[c].v = seg[7 + 3 * c] & 15;
                comps[c].tq = seg[8 + 3 * c];
                if (comps[c].h < 1 || comps[c].h > 2 || comps[c].v < 1 || comps[c].v > 2 || comps[c].tq > 3) throw std::runtime_error("jpeg: unsupported sampling factors");
                hmax = std::max(hmax, comps[c].h);
                vmax = std::max(vmax, comps[c].v);
            }
            if (nc == 1) comps[0].h = comps[0].v = hmax = vmax = 1;  // a single component is never interleaved
            const uint32_t mcux = (width + 8 * hmax - 1) / (8 * hmax), mcuy = (height + 8 * vmax - 1) / (8 * vmax);
            for (Comp &c : comps) {
                c.bw = mcux * c.h;
                c.bh = mcuy * c.v;
                c.rbw = ((width * c.h + hmax - 1) / hmax + 7) / 8;
                c.rbh = ((height * c.v + vmax - 1) / vmax + 7) / 8;
                c.coef.assign((size_t)c.bw * c.bh * 64, 0);
            }
            have_frame = true;
        } else if (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
            throw std::runtime_error("jpeg: lossless, hierarchical and arithmetic-coded files are not supported");
        } else if (m == 0xee) {  // APP14 "Adobe": colour transform flag (0 = components are R, G, B as they stand; 1 = YCbCr)
            if (n >= 12 && std::memcmp(seg, "Adobe", 5) == 0) adobe_transform = seg[11];
        } else if (m == 0xdd) {
            if (n < 2) throw std::runtime_error("jpeg: bad DRI");
            restart = (seg[0] << 8) | seg[1];
        } else if (m == 0xda) {  // SOS
            if (!have_frame) throw std::runtime_error("jpeg: SOS before SOF");
            const int ns = seg[0];
            if (ns < 1 || ns > (int)comps.size() || n < (size_t)(1 + 2 * ns + 3)) throw std::runtime_error("jpeg: bad SOS");
            std::vector<Comp *> sc;
            for (int k = 0; k < ns; ++k) {
                Comp *c = nullptr;
                for (Comp &cc : comps)
                    if (cc.id == seg[1 + 2 * k]) c = &cc;
                if (!c) throw std::runtime_error("jpeg: scan names an unknown component");
                c->td = seg[2 + 2 * k] >> 4;
                c->ta = seg[2 + 2 * k] & 15;
                if (c->td > 3 || c->ta > 3) throw std::runtime_error("jpeg: bad table selector");
                c->pred = 0;
                sc.push_back(c);
            }
            const int ss = seg[1 + 2 * ns], se = seg[2 + 2 * ns], ah = seg[3 + 2 * ns] >> 4, al = seg[3 + 2 * ns] & 15;
            if (progressive) {
                if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss != 0 && ns != 1) || al > 13) throw std::runtime_error("jpeg: bad progressive scan parameters");
            } else if (ss != 0 || se != 63 || ah != 0 || al != 0) {
                throw std::runtime_error("jpeg: bad sequential scan parameters");
            }
            for (Comp *c : sc) {
                const bool need_dc = !progressive || ss == 0, need_ac = !progressive || ss != 0;
                if ((need_dc && !(progressive && ah) && !dc[c->td].present) || (need_ac && !ac[c->ta].present)) throw std::runtime_error("jpeg: missing Huffman table");
            }
            JpegBits br{data + pos + seglen, data + len};
            uint32_t until_restart = restart, eobrun = 0;
            auto decode_block = [&](Comp &c, int16_t *blk) {
                if (!progressive) {
                    const int t = br.decode(dc[c.td]);
                    if (t > 11) throw std::runtime_error("jpeg: bad DC size");
                    c.pred += jpeg_extend(br.receive(t), t);
                    blk[0] = (int16_t)c.pred;
                    for (int k = 1; k < 64;) {
                        const int rs = br.decode(ac[c.ta]), r = rs >> 4, sz = rs & 15;
                        if (sz == 0) {
                            if (r != 15) break;
                            k += 16;
                            continue;
                        }
                        k += r;
                        if (k > 63) throw std::runtime_error("jpeg: bad AC run");
                        blk[zigzag[k]] = (int16_t)jpeg_extend(br.receive(sz), sz);
                        ++k;
                    }
                } else if (ss == 0) {  // DC: first pass or one more bit
                    if (ah == 0) {
                        const int t = br.decode(dc[c.td]);
                        if (t > 11) throw std::runtime_error("jpeg: bad DC size");
                        c.pred += jpeg_extend(br.receive(t), t);
                        blk[0] = (int16_t)(c.pred * (1 << al));
                    } else if (br.bit()) {
                        blk[0] = (int16_t)(blk[0] + (1 << al));
                    }
                } else if (ah == 0) {  // AC band, first pass
                    if (eobrun) {
                        --eobrun;
                        return;
                    }
                    for (int k = ss; k <= se;) {
                        const int rs = br.decode(ac[c.ta]), r = rs >> 4, sz = rs & 15;
                        if (sz == 0) {
                            if (r < 15) {
                                eobrun = (1u << r) - 1u;
                                if (r) eobrun += (uint32_t)br.receive(r);
                                break;
                            }
                            k += 16;
                        } else {
                            k += r;
                            if (k > 63) throw std::runtime_error("jpeg: bad AC run");
                            blk[zigzag[k]] = (int16_t)(jpeg_extend(br.receive(sz), sz) * (1 << al));
                            ++k;
                        }
                    }
                } else {  // AC band, refinement: one more bit for the non-zero coefficients, new +-1 coefficients in between
                    const int bit = 1 << al;
                    auto refine = [&](int16_t &p) {
                        if (br.bit() && (p & bit) == 0) p = (int16_t)(p > 0 ? p + bit : p - bit);
                    };
                    if (eobrun) {
                        --eobrun;
                        for (int k = ss; k <= se; ++k)
                            if (blk[zigzag[k]]) refine(blk[zigzag[k]]);
                        return;
                    }
                    for (int k = ss; k <= se;) {
                        const int rs = br.decode(ac[c.ta]), sz = rs & 15;
                        int r = rs >> 4, value = 0;
                        if (sz == 0) {
                            if (r < 15) {
                                eobrun = (1u << r) - 1u;
                                if (r) eobrun += (uint32_t)br.receive(r);
                                r = 64;  // run to the end of the band, refining on the way
                            }
                        } else {
                            if (sz != 1) throw std::runtime_error("jpeg: bad refinement code");
                            value = br.bit() ? bit : -bit;
                        }
                        while (k <= se) {
                            int16_t &p = blk[zigzag[k++]];
                            if (p) {
                                refine(p);
                            } else {
                                if (r == 0) {
                                    p = (int16_t)value;
                                    break;
                                }
                                --r;
                            }
                        }
                    }
                }
            };
            auto maybe_restart = [&]() {
                if (restart && until_restart == 0) {  // RSTn: byte-align, skip the marker, reset predictors and the EOB run
                    br.reset();
                    while (br.p + 1 < br.end && !(br.p[0] == 0xff && br.p[1] >= 0xd0 && br.p[1] <= 0xd7)) ++br.p;
                    if (br.p + 1 < br.end) br.p += 2;
                    for (Comp *c : sc) c->pred = 0;
                    eobrun = 0;
                    until_restart = restart;
                }
            };
            if (ns == 1) {  // non-interleaved: the component's own blocks, row by row
                Comp &c = *sc[0];
                for (uint32_t by = 0; by < c.rbh; ++by)
                    for (uint32_t bx = 0; bx < c.rbw; ++bx) {
                        maybe_restart();
                        decode_block(c, c.coef.data() + ((size_t)by * c.bw + bx) * 64);
                        if (restart) --until_restart;
                    }
            } else {
                const uint32_t mcux = comps[0].bw / comps[0].h, mcuy = comps[0].bh / comps[0].v;
                for (uint32_t my = 0; my < mcuy; ++my)
                    for (uint32_t mx = 0; mx < mcux; ++mx) {
                        maybe_restart();
                        for (Comp *c : sc)
                            for (int by = 0; by < c->v; ++by)
                                for (int bx = 0; bx < c->h; ++bx)
                                    decode_block(*c, c->coef.data() + ((size_t)(my * c->v + by) * c->bw + (mx * c->h + bx)) * 64);
                        if (restart) --until_restart;
                    }
            }
            decoded = true;
            pos = (size_t)(br.p - data);  // continue the marker search behind the entropy-coded segment
            continue;
        }
        pos += seglen;
    }
    if (decoded) {  // dequantise + IDCT every block
        int deq[64];
        for (Comp &c : comps) {
            if (!have_qt[c.tq]) throw std::runtime_error("jpeg: missing quantisation table");
            const size_t stride = (size_t)c.bw * 8;
            c.plane.assign(stride * c.bh * 8, 0);
            for (uint32_t by = 0; by < c.bh; ++by)
                for (uint32_t bx = 0; bx < c.bw; ++bx) {
                    const int16_t *blk = c.coef.data() + ((size_t)by * c.bw + bx) * 64;
                    for (int k = 0; k < 64; ++k) {  // (a real coefficient stays below 2^19; the clamp only tames corrupt tables)
                        const int64_t v = (int64_t)blk[k] * qt[c.tq][k];
                        deq[k] = (int)(v < -(1 << 24) ? -(1 << 24) : (v > (1 << 24) ? (1 << 24) : v));
                    }
                    jpeg_idct(deq, c.plane.data() + (size_t)by * 8 * stride + (size_t)bx * 8, stride);
                }
        }
    }
    if (!decoded) throw std::runtime_error("jpeg: no image data");
    // upsample the chroma planes to full resolution (jpeg-decoder's upsampler.rs: triangle filters for the factor-2 cases)
    auto upsample = [&](const Comp &c, std::vector<uint8_t> &out) {
        const uint32_t cw = (width * c.h + hmax - 1) / hmax, chh = (height * c.v + vmax - 1) / vmax;  // the component's own size
        const size_t stride = (size_t)c.bw * 8;
        const bool h2 = hmax == 2 && c.h == 1, v2 = vmax == 2 && c.v == 1;
        out.assign((size_t)width * height, 0);
        std::vector<uint8_t> line((size_t)cw * 2 + 2);
        for (uint32_t y = 0; y < height; ++y) {
            const uint8_t *near_row, *far_row;
            if (v2) {
                const uint32_t rn = std::min(y / 2, chh - 1);
                const uint32_t rf = (y & 1) ? std::min(rn + 1, chh - 1) : (rn ? rn - 1 : 0);
                near_row = c.plane.data() + (size_t)rn * stride;
                far_row = c.plane.data() + (size_t)rf * stride;
            } else {
                near_row = far_row = c.plane.data() + (size_t)std::min(y, chh - 1) * stride;
            }
            uint8_t *o = out.data() + (size_t)y * width;
            if (!h2 && !v2) {
                std::memcpy(o, near_row, width);
            } else if (!h2) {  // h1v2
                for (uint32_t x = 0; x < width; ++x) o[x] = (uint8_t)((3 * near_row[x] + far_row[x] + 2) >> 2);
            } else if (!v2) {  // h2v1
                if (cw == 1) {
                    line[0] = line[1] = near_row[0];
                } else {
                    line[0] = near_row[0];
                    line[1] = (uint8_t)((near_row[0] * 3 + near_row[1] + 2) >> 2);
                    for (uint32_t i = 1; i + 1 < cw; ++i) {
                        const int s3 = 3 * near_row[i] + 2;
                        line[2 * i] = (uint8_t)((s3 + near_row[i - 1]) >> 2);
                        line[2 * i + 1] = (uint8_t)((s3 + near_row[i + 1]) >> 2);
                    }
                    line[2 * (cw - 1)] = (uint8_t)((near_row[cw - 1] * 3 + near_row[cw - 2] + 2) >> 2);
                    line[2 * (cw - 1) + 1] = near_row[cw - 1];
                }
                std::memcpy(o, line.data(), width);
            } else {  // h2v2
                if (cw == 1) {
                    line[0] = line[1] = (uint8_t)((3 * near_row[0] + far_row[0] + 2) >> 2);
                } else {
                    int t0 = 3 * near_row[0] + far_row[0], t1 = 3 * near_row[1] + far_row[1];
                    line[0] = (uint8_t)((t0 * 4 + 8) >> 4);
                    line[1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
                    for (uint32_t i = 2; i < cw; ++i) {
                        const int t2 = 3 * near_row[i] + far_row[i];
                        line[2 * i - 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
                        line[2 * i - 1] = (uint8_t)((3 * t1 + t2 + 8) >> 4);
                        t0 = t1;
                        t1 = t2;
                    }
                    line[2 * cw - 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
                    line[2 * cw - 1] = (uint8_t)((t1 * 4 + 8) >> 4);
                }
                std::memcpy(o, line.data(), width);
            }
        }
    };
    std::vector<uint8_t> full[3];
    for (size_t c = 0; c < comps.size(); ++c) upsample(comps[c], full[c]);
    rgba8.resize((size_t)width * height * 4);
    auto f2f20 = [](float x) { return (int)(x * (float)(1 << 20) + 0.5f); };
    for (uint32_t y = 0; y < height; ++y) {
        const uint32_t fy = height - 1 - y;  // DynamicImage::flipv
        for (uint32_t x = 0; x < width; ++x) {
            const size_t i = (size_t)y * width + x;
            uint8_t r, g, b;
            if (comps.size() == 1) {
                r = g = b = full[0][i];
            } else if (adobe_transform == 0) {  // Adobe RGB JPEG: no colour transform
                r = full[0][i];
                g = full[1][i];
                b = full[2][i];
            } else {  // BT.601, 20-bit fixed point
                const int yy = ((int)full[0][i] << 20) + (1 << 19), cb = (int)full[1][i] - 128, cr = (int)full[2][i] - 128;
                r = jpeg_clamp((yy + f2f20(1.40200f) * cr) >> 20);
                g = jpeg_clamp((yy - f2f20(0.34414f) * cb - f2f20(0.71414f) * cr) >> 20);
                b = jpeg_clamp((yy + f2f20(1.77200f) * cb) >> 20);
            }
            uint8_t *o = rgba8.data() + ((size_t)fy * width + x) * 4;
            o[0] = r; o[1] = g; o[2] = b; o[3] = 255;
        }
    }
}

// ---- TIFF (the `tiff` 0.9 decoder behind image::ImageFormat::Tiff -> flipv -> to_rgba8, load.rs:590-603) ----------------
// Baseline strips, chunky layout, 8 / 16 bits per sample, gray / gray + alpha / RGB / RGBA, compression none / LZW / Deflate /
// PackBits, horizontal predictor, either byte order.  Tiles, planar layout, palettes, CMYK / YCbCr and fax codecs are rejected.
void tiff_lzw(const uint8_t *src, size_t n, std::vector<uint8_t> &out, size_t expect) {
    std::vector<uint32_t> prefix(4096), length(4096);
    std::vector<uint8_t> suffix(4096), first(4096);
    for (uint32_t i = 0; i < 256; ++i) {
        prefix[i] = 0xffffu;
        length[i] = 1;
        suffix[i] = first[i] = (uint8_t)i;
    }
    uint32_t next = 258, bits = 9, old = 0xffffu;
    uint64_t acc = 0;
    int nacc = 0;
    size_t pos = 0;
    std::vector<uint8_t> tmp;
    auto emit = [&](uint32_t code) {
        const uint32_t len = length[code];
        const size_t base = out.size();
        out.resize(base + len);
        uint32_t c = code;
        for (uint32_t k = len; k-- > 0;) {
            out[base + k] = suffix[c];
            c = prefix[c];
        }
    };
    while (out.size() < expect) {
        while (nacc < (int)bits && pos < n) {
            acc = (acc << 8) | src[pos++];
            nacc += 8;
        }
        if (nacc < (int)bits) break;
        const uint32_t code = (uint32_t)(acc >> (nacc - bits)) & ((1u << bits) - 1u);
        nacc -= bits;
        if (code == 257u) break;  // EOI
        if (code == 256u) {       // clear
            next = 258;
            bits = 9;
            old = 0xffffu;
            continue;
        }
        if (old == 0xffffu) {
            if (code >= 256u) throw std::runtime_error("tiff: bad LZW stream");
            emit(code);
        } else if (code < next) {
            emit(code);
            if (next < 4096u) {
                prefix[next] = old;
                length[next] = length[old] + 1;
                suffix[next] = first[code];
                first[next] = first[old];
                ++next;
            }
        } else if (code == next && next < 4096u) {
            prefix[next] = old;
            length[next] = length[old] + 1;
            suffix[next] = first[old];
            first[next] = first[old];
            ++next;
            emit(code);
        } else {
            throw std::runtime_error("tiff: bad LZW code");
        }
        old = code;
        if (next + 1 >= (1u << bits) && bits < 12) ++bits;  // TIFF's early change: widen one code early
    }
}
void decode_tiff(const uint8_t *data, size_t len, uint32_t &width, uint32_t &height, std::vector<uint8_t> &rgba8) {
    if (len < 8) throw std::runtime_error("tiff: truncated file");
    const bool le = data[0] == 'I' && data[1] == 'I', be = data[0] == 'M' && data[1] == 'M';
    if (!le && !be) throw std::runtime_error("tiff: bad byte-order mark");
    auto need = [&](size_t pos, size_t n) {
        if (pos + n > len) throw std::runtime_error("tiff: truncated file");
    };
    auto r16 = [&](size_t p) -> uint32_t {
        need(p, 2);
        return le ? (uint32_t)data[p] | ((uint32_t)data[p + 1] << 8) : ((uint32_t)data[p] << 8) | data[p + 1];
    };
    auto r32 = [&](size_t p) -> uint32_t {
        need(p, 4);
        return le ? (uint32_t)data[p] | ((uint32_t)data[p + 1] << 8) | ((uint32_t)data[p + 2] << 16) | ((uint32_t)data[p + 3] << 24)
                  : ((uint32_t)data[p] << 24) | ((uint32_t)data[p + 1] << 16) | ((uint32_t)data[p + 2] << 8) | data[p + 3];
    };
    if (r16(2) != 42u) throw std::runtime_error("tiff: bad magic number (BigTIFF is not supported)");
    const size_t ifd = r32(4);
    const uint32_t n_entries = r16(ifd);
    uint32_t compression = 1, photometric = 1, spp = 1, rows_per_strip = 0xffffffffu, planar = 1, predictor = 1, extra = 0;
    std::vector<uint32_t> bits, strip_offsets, strip_counts;
    width = height = 0;
    auto values = [&](size_t entry, std::vector<uint32_t> &out) {  // SHORT / LONG arrays, inline when they fit in 4 bytes
        const uint32_t type = r16(entry + 2), count = r32(entry + 4);
        const uint32_t size = type == 3 ? 2 : (type == 4 ? 4 : (type == 1 ? 1 : 0));
        if (!size) throw std::runtime_error("tiff: unexpected field type");
        if ((uint64_t)size * count > len) throw std::runtime_error("tiff: field longer than the file");
        size_t p = (size_t)size * count <= 4 ? entry + 8 : r32(entry + 8);
        out.resize(count);
        for (uint32_t i = 0; i < count; ++i, p += size) out[i] = size == 2 ? r16(p) : (size == 4 ? r32(p) : (need(p, 1), data[p]));
    };
    for (uint32_t e = 0; e < n_entries; ++e) {
        const size_t entry = ifd + 2 + (size_t)e * 12;
        need(entry, 12);
        std::vector<uint32_t> v;
        switch (r16(entry)) {
        case 256: values(entry, v); width = v.at(0); break;
        case 257: values(entry, v); height = v.at(0); break;
        case 258: values(entry, bits); break;
        case 259: values(entry, v); compression = v.at(0); break;
        case 262: values(entry, v); photometric = v.at(0); break;
        case 273: values(entry, strip_offsets); break;
        case 277: values(entry, v); spp = v.at(0); break;
        case 278: values(entry, v); rows_per_strip = v.at(0); break;
        case 279: values(entry, strip_counts); break;
        case 284: values(entry, v); planar = v.at(0); break;
        case 317: values(entry, v); predictor = v.at(0); break;
        case 338: values(entry, v); extra = (uint32_t)v.size(); break;
        case 322: case 323: case 324: case 325: throw std::runtime_error("tiff: tiled files are not supported");
        default: break;
        }
    }
    if (!width || !height || strip_offsets.empty() || strip_offsets.size() != strip_counts.size()) throw std::runtime_error("tiff: missing size or strips");
    check_image_size(width, height, len, "tiff");
    if (bits.empty()) bits.assign(1, 1);
    const uint32_t bps = bits[0];
    for (uint32_t b : bits)
        if (b != bps) throw std::runtime_error("tiff: mixed bits per sample");
    if (bps != 8 && bps != 16) throw std::runtime_error("tiff: only 8 / 16 bits per sample are supported");
    if (planar != 1 && spp > 1) throw std::runtime_error("tiff: planar layout is not supported");
    if (photometric > 2 || (photometric == 2 && spp < 3) || spp > 4 || spp == 0) throw std::runtime_error("tiff: unsupported photometric interpretation / sample count");
    (void)extra;
    const size_t bpp = (size_t)spp * bps / 8, stride = (size_t)width * bpp;
    rows_per_strip = std::min(rows_per_strip, height);
    std::vector<uint8_t> img;
    img.reserve(stride * height);
    for (size_t sidx = 0; sidx < strip_offsets.size() && img.size() < stride * height; ++sidx) {
        const size_t rows = std::min<size_t>(rows_per_strip, height - sidx * rows_per_strip), expect = rows * stride;
        need(strip_offsets[sidx], strip_counts[sidx]);
        const uint8_t *src = data + strip_offsets[sidx];
        const size_t n = strip_counts[sidx], base = img.size();
        std::vector<uint8_t> strip;
        if (compression == 1) {
            if (n < expect) throw std::runtime_error("tiff: short strip");
            strip.assign(src, src + expect);
        } else if (compression == 5) {
            tiff_lzw(src, n, strip, expect);
        } else if (compression == 8 || compression == 32946) {
            strip.resize(expect);
            uLongf out_len = static_cast<uLongf>(expect);
            if (uncompress(strip.data(), &out_len, src, static_cast<uLong>(n)) != Z_OK) throw std::runtime_error("tiff: inflate failed");
            strip.resize(out_len);
        } else if (compression == 32773) {  // PackBits
            size_t i = 0;
            while (i < n && strip.size() < expect) {
                const int c = (int8_t)src[i++];
                if (c >= 0) {
                    if (i + (size_t)c + 1 > n) throw std::runtime_error("tiff: bad PackBits data");
                    strip.insert(strip.end(), src + i, src + i + c + 1);
                    i += (size_t)c + 1;
                } else if (c != -128) {
                    if (i >= n) throw std::runtime_error("tiff: bad PackBits data");
                    strip.insert(strip.end(), (size_t)(1 - c), src[i++]);
                }
            }
        } else {
            throw std::runtime_error("tiff: compression " + std::to_string(compression) + " is not supported (none, LZW, Deflate, PackBits are)");
        }
        if (strip.size() < expect) throw std::runtime_error("tiff: strip decodes to too few bytes");
        img.insert(img.end(), strip.begin(), strip.begin() + expect);
        if (predictor == 2) {  // horizontal differencing, per sample, in the file's byte order for 16-bit samples
            for (size_t r = 0; r < rows; ++r) {
                uint8_t *row = img.data() + base + r * stride;
                if (bps == 8) {
                    for (size_t x = bpp; x < stride; ++x) row[x] = static_cast<uint8_t>(row[x] + row[x - bpp]);
                } else {
                    for (size_t x = bpp; x + 1 < stride + 1 && x < stride; x += 2) {
                        const uint32_t a = le ? row[x - bpp] | (row[x - bpp + 1] << 8) : (row[x - bpp] << 8) | row[x - bpp + 1];
                        const uint32_t b = le ? row[x] | (row[x + 1] << 8) : (row[x] << 8) | row[x + 1];
                        const uint32_t v = (a + b) & 0xffffu;
                        if (le) {
                            row[x] = (uint8_t)v;
                            row[x + 1] = (uint8_t)(v >> 8);
                        } else {
                            row[x] = (uint8_t)(v >> 8);
                            row[x + 1] = (uint8_t)v;
                        }
                    }
                }
            }
        } else if (predictor != 1) {
            throw std::runtime_error("tiff: floating-point predictor is not supported");
        }
    }
    if (img.size() < stride * height) throw std::runtime_error("tiff: image data ends early");
    rgba8.resize((size_t)width * height * 4);
    for (uint32_t y = 0; y < height; ++y) {
        const uint32_t fy = height - 1 - y;  // DynamicImage::flipv
        for (uint32_t x = 0; x < width; ++x) {
            const uint8_t *px = img.data() + y * stride + x * bpp;
            auto sample = [&](uint32_t c) -> uint8_t {
                if (bps == 8) return px[c];
                const uint32_t v = le ? px[2 * c] | (px[2 * c + 1] << 8) : (px[2 * c] << 8) | px[2 * c + 1];
                return static_cast<uint8_t>((v + 128u) / 257u);  // to_rgba8 (see decode_png)
            };
            uint8_t r, g, b, a = 255;
            if (photometric <= 1) {
                uint8_t l = sample(0);
                if (photometric == 0) l = static_cast<uint8_t>(255 - l);  // WhiteIsZero
                r = g = b = l;
                if (spp >= 2) a = sample(1);
            } else {
                r = sample(0);
                g = sample(1);
                b = sample(2);
                if (spp == 4) a = sample(3);
            }
            uint8_t *o = rgba8.data() + ((size_t)fy * width + x) * 4;
            o[0] = r; o[1] = g; o[2] = b; o[3] = a;
        }
    }
}

// ---- OpenEXR (the `image` crate's OpenExrDecoder -> to_rgba32f, load.rs:590-610) -----------------------------------------
// Single-part scanline files with R, G, B (and optionally A) channels of type HALF or FLOAT, compression NONE / RLE / ZIPS /
// ZIP (lossless, bit-exact by construction).  Tiled, multi-part and deep files and the PIZ / PXR24 / B44 / DWA codecs are
// rejected.  Rows are flipped (DynamicImage::flipv); a missing alpha channel reads 1.
float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {  // subnormal half: normalise
            int e = -1;
            uint32_t m = man;
            do {
                ++e;
                m <<= 1;
            } while (!(m & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((m & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + (127 - 15)) << 23) | (man << 13);
    }
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}
void decode_exr(const uint8_t *data, size_t len, uint32_t &width, uint32_t &height, std::vector<float> &rgba) {
    auto need = [&](size_t pos, size_t n) {
        if (pos + n > len) throw std::runtime_error("exr: truncated file");
    };
    auto rd32 = [&](size_t pos) {
        need(pos, 4);
        uint32_t v;
        std::memcpy(&v, data + pos, 4);
        return v;
    };
    if (len < 8 || rd32(0) != 20000630u) throw std::runtime_error("exr: bad magic number");
    const uint32_t version = rd32(4);
    if ((version & 0xffu) != 2u) throw std::runtime_error("exr: unsupported file version");
    if (version & 0x1a00u) throw std::runtime_error("exr: tiled, deep and multi-part files are not supported");
    struct Channel {
        std::string name;
        uint32_t type;  // 0 UINT, 1 HALF, 2 FLOAT
    };
    std::vector<Channel> channels;
    int32_t win[4] = {0, 0, -1, -1};
    uint32_t compression = 0xffu;
    size_t pos = 8;
    while (true) {  // attributes: name\0 type\0 size data
        need(pos, 1);
        if (data[pos] == 0) {
            ++pos;
            break;
        }
        auto cstr = [&]() {
            size_t e = pos;
            while (e < len && data[e]) ++e;
            if (e >= len) throw std::runtime_error("exr: truncated header");
            std::string r(reinterpret_cast<const char *>(data + pos), e - pos);
            pos = e + 1;
            return r;
        };
        const std::string name = cstr(), type = cstr();
        const uint32_t size = rd32(pos);
        pos += 4;
        need(pos, size);
        if (name == "channels" && type == "chlist") {
            size_t q = pos;
            while (q < pos + size && data[q]) {
                size_t e = q;
                while (e < pos + size && data[e]) ++e;
                Channel c;
                c.name.assign(reinterpret_cast<const char *>(data + q), e - q);
                q = e + 1;
                if (q + 16 > pos + size) throw std::runtime_error("exr: truncated channel list");
                c.type = rd32(q);
                if (rd32(q + 8) != 1u || rd32(q + 12) != 1u) throw std::runtime_error("exr: subsampled channels are not supported");
                q += 16;
                channels.push_back(c);
            }
        } else if (name == "compression") {
            compression = data[pos];
        } else if (name == "dataWindow") {
            if (size != 16) throw std::runtime_error("exr: bad dataWindow");
            std::memcpy(win, data + pos, 16);
        }
        pos += size;
    }
    if (channels.empty() || win[2] < win[0] || win[3] < win[1]) throw std::runtime_error("exr: missing channels or dataWindow");
    uint32_t lines_per_chunk;
    switch (compression) {
    case 0: case 1: case 2: lines_per_chunk = 1; break;  // NONE, RLE, ZIPS
    case 3: lines_per_chunk = 16; break;                 // ZIP
    default: throw std::runtime_error("exr: compression " + std::to_string(compression) + " is not supported (NONE, RLE, ZIPS, ZIP are)");
    }
    if ((int64_t)win[2] - win[0] >= 65535 || (int64_t)win[3] - win[1] >= 65535) throw std::runtime_error("exr: data window out of range");
    width = (uint32_t)(win[2] - win[0] + 1);
    height = (uint32_t)(win[3] - win[1] + 1);
    check_image_size(width, height, len, "exr");
    int slot[4] = {-1, -1, -1, -1};  // index into `channels` of R, G, B, A
    size_t line_bytes = 0;
    std::vector<size_t> ch_offset(channels.size());
    for (size_t i = 0; i < channels.size(); ++i) {
        if (channels[i].type != 1u && channels[i].type != 2u) throw std::runtime_error("exr: only HALF and FLOAT channels are supported");
        ch_offset[i] = line_bytes;
        line_bytes += (size_t)width * (channels[i].type == 1u ? 2 : 4);
        static const char *names[4] = {"R", "G", "B", "A"};
        for (int k = 0; k < 4; ++k)
            if (channels[i].name == names[k]) slot[k] = (int)i;
    }
    if (slot[0] < 0 || slot[1] < 0 || slot[2] < 0) throw std::runtime_error("exr: image contains no R, G, B channels");
    const uint32_t n_chunks = (height + lines_per_chunk - 1) / lines_per_chunk;
    need(pos, (size_t)n_chunks * 8);
    rgba.assign((size_t)width * height * 4, 1.0f);
    std::vector<uint8_t> buf, tmp;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        uint64_t off;
        std::memcpy(&off, data + pos + (size_t)c * 8, 8);
        need(off, 8);
        const int32_t y0 = (int32_t)rd32(off) - win[1];
        const uint32_t csize = rd32(off + 4);
        need(off + 8, csize);
        if (y0 < 0 || (uint32_t)y0 >= height) throw std::runtime_error("exr: chunk outside the data window");
        const uint32_t lines = std::min(lines_per_chunk, height - (uint32_t)y0);
        const size_t raw_size = line_bytes * lines;
        const uint8_t *src = data + off + 8;
        buf.resize(raw_size);
        if (compression == 0 || csize == raw_size) {  // stored (a codec falls back to raw bytes when it does not shrink the chunk)
            if (csize != raw_size) throw std::runtime_error("exr: bad chunk size");
            std::memcpy(buf.data(), src, raw_size);
        } else {
            tmp.resize(raw_size);
            if (compression == 1) {  // RLE: n >= 0: one byte repeated n + 1 times; n < 0: -n literal bytes
                size_t i = 0, o = 0;
                while (i < csize) {
                    const int n = (int8_t)src[i++];
                    if (n < 0) {
                        const size_t k = (size_t)(-n);
                        if (i + k > csize || o + k > raw_size) throw std::runtime_error("exr: bad RLE data");
                        std::memcpy(tmp.data() + o, src + i, k);
                        i += k;
                        o += k;
                    } else {
                        const size_t k = (size_t)n + 1;
                        if (i >= csize || o + k > raw_size) throw std::runtime_error("exr: bad RLE data");
                        std::memset(tmp.data() + o, src[i++], k);
                        o += k;
                    }
                }
                if (o != raw_size) throw std::runtime_error("exr: bad RLE data");
            } else {
                uLongf out_len = static_cast<uLongf>(raw_size);
                if (uncompress(tmp.data(), &out_len, src, csize) != Z_OK || out_len != raw_size) throw std::runtime_error("exr: inflate failed");
            }
            for (size_t i = 1; i < raw_size; ++i) tmp[i] = static_cast<uint8_t>(tmp[i - 1] + tmp[i] - 128);  // undo the byte predictor
            const size_t half = (raw_size + 1) / 2;                                                          // then the even / odd byte split
            for (size_t i = 0; i < raw_size; ++i) buf[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
        }
        for (uint32_t l = 0; l < lines; ++l) {
            const uint32_t fy = height - 1 - ((uint32_t)y0 + l);  // DynamicImage::flipv
            const uint8_t *line = buf.data() + (size_t)l * line_bytes;
            for (int k = 0; k < 4; ++k) {
                if (slot[k] < 0) continue;
                const uint8_t *chp = line + ch_offset[(size_t)slot[k]];
                const bool is_half = channels[(size_t)slot[k]].type == 1u;
                for (uint32_t x = 0; x < width; ++x) {
                    float v;
                    if (is_half) {
                        uint16_t h;
                        std::memcpy(&h, chp + 2 * x, 2);
                        v = half_to_float(h);
                    } else {
                        std::memcpy(&v, chp + 4 * x, 4);
                    }
                    rgba[((size_t)fy * width + x) * 4 + (size_t)k] = v;
                }
            }
        }
    }
}

void decode_image(const std::string &format, const uint8_t *bytes, size_t len, uint32_t width, uint32_t height, uint32_t channels, AkrImage &img,
                  std::vector<uint8_t> &texels) {
    if (channels == 0 || channels > 4) throw std::runtime_error("Invalid number of channels: " + std::to_string(channels));
    if (format == "float") {
        if (len != (size_t)width * height * channels * 4) throw std::runtime_error("float image: buffer length does not match width * height * channels * 4");
        img.texel_format = AKR_TEXEL_RGBA32F;
        img.width = width;
        img.height = height;
        std::vector<float> rgba((size_t)width * height * 4);
        const float *src = reinterpret_cast<const float *>(bytes);
        std::vector<float> tmp((size_t)width * height * channels);
        std::memcpy(tmp.data(), src, len);
        for (size_t i = 0; i < (size_t)width * height; ++i)
            for (uint32_t c = 0; c < 4; ++c) rgba[i * 4 + c] = c < channels ? tmp[i * channels + c] : (c == 3 ? 1.0f : 0.0f);
        texels.resize(rgba.size() * 4);
        std::memcpy(texels.data(), rgba.data(), texels.size());
    } else if (format == "png") {
        img.texel_format = AKR_TEXEL_RGBA8;
        decode_png(bytes, len, img.width, img.height, texels);
    } else if (format == "jpeg") {
        img.texel_format = AKR_TEXEL_RGBA8;
        decode_jpeg(bytes, len, img.width, img.height, texels);
    } else if (format == "tiff") {
        img.texel_format = AKR_TEXEL_RGBA8;
        decode_tiff(bytes, len, img.width, img.height, texels);
    } else if (format == "exr") {
        img.texel_format = AKR_TEXEL_RGBA32F;
        std::vector<float> rgba;
        decode_exr(bytes, len, img.width, img.height, rgba);
        texels.resize(rgba.size() * 4);
        std::memcpy(texels.data(), rgba.data(), texels.size());
    } else {
        throw std::runtime_error("image format '" + format + "' needs a decoder this host does not carry (png, jpeg, tiff, exr and raw float are implemented)");
    }
}

class ShaderCompiler {
  public:
    ShaderCompiler(const akr::json::Value &graph, ImageResolver images) : graph_(graph), nodes_(graph.at("nodes")), images_(std::move(images)) {}
    CompiledShader compile() {
        const std::string &kind = graph_.at("kind").string();
        if (kind != "surface") throw std::runtime_error("shader kind '" + kind + "' is not supported (compiler.rs:277-286)");
        compile_node(graph_.at("output").at("id").string());
        CompiledShader out;
        out.nodes = std::move(bytecode_);
        out.data = std::move(data_);
        return out;
    }

  private:
    const akr::json::Value &graph_;
    const akr::json::Value &nodes_;
    ImageResolver images_;
    std::map<std::string, uint32_t> env_;
    std::set<std::string> in_progress_;  // nodes on the current compile_node recursion path
    std::vector<AkrSvmNode> bytecode_;
    std::vector<uint8_t> data_;

    // util::ByteVecBuilder::push (util/mod.rs:481-494): offset rounded up to align_of::<T>().
    uint32_t push_bytes(const void *p, size_t size, size_t align) {
        size_t off = (data_.size() + align - 1) / align * align;
        data_.resize(off + size, 0);
        std::memcpy(data_.data() + off, p, size);
        return static_cast<uint32_t>(off);
    }
    uint32_t push_f32(float v) { return push_bytes(&v, 4, 4); }
    uint32_t push_u32(uint32_t v) { return push_bytes(&v, 4, 4); }
    uint32_t opt_ref(const akr::json::Value &node, const char *field) {
        if (!node.has(field) || node.at(field).is_null()) return AKR_SVM_NONE;
        return ref(node, field);
    }
    uint32_t push_float3(const float v[3]) {
        // luisa Float3: 16-byte size and alignment; the pad lane is zero.
        float q[4] = {v[0], v[1], v[2], 0.0f};
        return push_bytes(q, 16, 16);
    }
    uint32_t push(uint32_t op, std::initializer_list<uint32_t> args) {
        AkrSvmNode n{};
        n.op = op;
        n.n_args = static_cast<uint32_t>(args.size());
        uint32_t k = 0;
        for (uint32_t a : args) n.a[k++] = a;
        bytecode_.push_back(n);
        return static_cast<uint32_t>(bytecode_.size() - 1);
    }
    uint32_t ref(const akr::json::Value &node, const char *field) {
        return compile_node(node.at(field).at("id").string());
    }
    uint32_t compile_node(const std::string &id) {
        auto it = env_.find(id);
        if (it != env_.end()) return it->second;
        // a node that (transitively) feeds itself: the reference's compiler recurses until its stack overflows
        // (compiler.rs:116-337 has no visited set either); rejected here
        if (!in_progress_.insert(id).second) throw std::runtime_error("shader graph has a cycle through node '" + id + "'");
        struct Done {
            std::set<std::string> &s;
            const std::string &id;
            ~Done() { s.erase(id); }
        } done{in_progress_, id};
        const akr::json::Value &node = nodes_.at(id);
        const std::string &type = node.at("type").string();
        uint32_t idx;
        if (type == "float") {
            idx = push(AKR_SVM_FLOAT, {push_f32(node.at("value").f32())});
        } else if (type == "float3") {
            float v[3] = {node.at("value").at(0).f32(), node.at("value").at(1).f32(), node.at("value").at(2).f32()};
            idx = push(AKR_SVM_FLOAT3, {push_float3(v)});
        } else if (type == "rgb") {
            float v[3] = {node.at("value").at(0).f32(), node.at("value").at(1).f32(), node.at("value").at(2).f32()};
            const std::string &cs = node.at("colorspace").string();
            uint32_t cs_id;
            if (cs == "srgb") cs_id = 1;          // ColorSpaceId::SRGB (color.rs:29-47)
            else if (cs == "aces") cs_id = 2;
            else throw std::runtime_error("rgb node: unknown colorspace '" + cs + "'");
            uint32_t data = push(AKR_SVM_FLOAT3, {push_float3(v)});
            idx = push(AKR_SVM_RGB_TEX, {data, cs_id});
        } else if (type == "spectral_uplift") {
            idx = push(AKR_SVM_SPECTRAL_UPLIFT, {ref(node, "rgb")});
        } else if (type == "diffuse") {
            idx = push(AKR_SVM_DIFFUSE_BSDF, {ref(node, "color")});
        } else if (type == "emission") {
            uint32_t c = ref(node, "color");
            uint32_t s = ref(node, "strength");
            idx = push(AKR_SVM_EMISSION, {c, s});
        } else if (type == "glass") {
            uint32_t c = ref(node, "color");
            uint32_t ior = ref(node, "ior");
            uint32_t rough = ref(node, "roughness");
            idx = push(AKR_SVM_GLASS_BSDF, {c, c, rough, ior});
        } else if (type == "principled") {
            // compile order = field order in compiler.rs:181-205 (subsurface_ior is skipped)
            static const char *fields[25] = {
                "base_color", "metallic", "roughness", "ior", "alpha", "normal", "subsurface_weight",
                "subsurface_radius", "subsurface_scale", "subsurface_anisotropy", "specular_ior_level",
                "specular_tint", "anisotropic", "anisotropic_rotation", "tangent", "transmission_weight",
                "sheen_weight", "sheen_tint", "coat_weight", "coat_roughness", "coat_ior", "coat_tint",
                "coat_normal", "emission_color", "emission_strength"};
            AkrSvmNode n{};
            n.op = AKR_SVM_PRINCIPLED_BSDF;
            n.n_args = 25;
            for (int k = 0; k < 25; ++k) n.a[k] = ref(node, fields[k]);
            bytecode_.push_back(n);
            idx = static_cast<uint32_t>(bytecode_.size() - 1);
        } else if (type == "output") {
            idx = push(AKR_SVM_MATERIAL_OUTPUT, {ref(node, "node")});
        } else if (type == "image") {  // compiler.rs:135-160
            const akr::json::Value &image = node.at("image");
            const std::string &cs = image.at("colorspace").string();
            uint32_t cs_id;
            if (cs == "none") cs_id = 0;  // ColorSpaceId::NONE: no gamma decode
            else if (cs == "srgb") cs_id = 1;
            else throw std::runtime_error("image node: colorspace '" + cs + "' is todo!() in the reference (texture/mod.rs:52-58)");
            uint32_t tex = push_u32(images_(image));
            uint32_t uv = opt_ref(node, "uv");
            idx = push(AKR_SVM_RGB_IMAGE_TEX, {tex, cs_id, uv});
        } else if (type == "texcoords") {
            idx = push(AKR_SVM_TEX_COORDS, {});
        } else if (type == "extract") {  // compiler.rs:272-276
            uint32_t n = ref(node, "node");
            const std::string &f = node.at("field").string();
            uint32_t fid;
            if (f == "uv") fid = AKR_SVM_FIELD_UV;
            else if (f == "Red") fid = AKR_SVM_FIELD_RED;
            else if (f == "Green") fid = AKR_SVM_FIELD_GREEN;
            else if (f == "Blue") fid = AKR_SVM_FIELD_BLUE;
            else throw std::runtime_error("extract node: unknown field '" + f + "'");
            idx = push(AKR_SVM_EXTRACT_FIELD, {n, fid});
        } else if (type == "mapping") {  // compiler.rs:288-304
            uint32_t v = ref(node, "vector");
            uint32_t loc = ref(node, "location");
            uint32_t rot = ref(node, "rotation");
            uint32_t sc = ref(node, "scale");
            const std::string &m = node.at("mapping").string();
            uint32_t ty;
            if (m == "point") ty = 0;
            else if (m == "texture") ty = 1;
            else throw std::runtime_error("mapping node: unknown type '" + m + "'");
            idx = push(AKR_SVM_MAPPING, {v, ty, loc, rot, sc});
        } else if (type == "normal_map") {  // compiler.rs:305-317
            uint32_t nrm = ref(node, "normal");
            uint32_t st = ref(node, "strength");
            if (node.at("space").string() != "tangent") throw std::runtime_error("normal_map: only tangent space is implemented in the reference (eval.rs:190-194)");
            idx = push(AKR_SVM_NORMAL_MAP, {nrm, st});
        } else if (type == "checkerboard") {  // compiler.rs:318-334
            uint32_t v = opt_ref(node, "vector");
            uint32_t sc = ref(node, "scale");
            uint32_t c1 = ref(node, "color1");
            uint32_t c2 = ref(node, "color2");
            idx = push(AKR_SVM_CHECKERBOARD, {v, sc, c1, c2});
        } else if (type == "separate_color") {  // compiler.rs:335-341
            if (node.at("mode").string() != "rgb") throw std::runtime_error("separate_color: unknown mode");
            idx = push(AKR_SVM_SEPARATE_COLOR, {ref(node, "color")});
        } else {
            // float4 / noise / mix / math / metal / plastic are todo!() (or unreachable) in the reference compiler
            throw std::runtime_error("shader node type '" + type + "' is todo!() in the reference compiler (svm/compiler.rs:134,161-163,262-266,342)");
        }
        env_[id] = idx;
        return idx;
    }
};

}  // namespace

// ---- AkrHostScene ----------------------------------------------------------------------------
struct AkrHostScene {
    std::map<std::string, std::vector<uint8_t>> buffers;
    struct MeshStore {
        std::vector<float> vertices, normals, uvs, tangents;
        std::vector<uint32_t> indices, material_slots;
        bool has_normals = false, has_uvs = false, has_tangents = false;
    };
    std::vector<MeshStore> mesh_store;
    std::vector<AkrMesh> meshes;
    std::vector<std::vector<AkrSvmNode>> kind_nodes;
    std::vector<AkrShaderKind> kinds;
    std::vector<uint8_t> shader_data;
    std::vector<std::vector<AkrShaderRef>> instance_materials;
    std::vector<AkrInstance> instances;
    std::vector<std::string> instance_names, geometry_names, material_names;
    std::vector<std::vector<uint8_t>> image_texels;  // decoded RGBA8 / RGBA32F texels, one per (image, sampler) slot
    std::vector<AkrImage> images;
    AkrSceneDesc desc{};

    void refresh_desc() {
        meshes.resize(mesh_store.size());
        for (size_t i = 0; i < mesh_store.size(); ++i) {
            MeshStore &m = mesh_store[i];
            AkrMesh &d = meshes[i];
            d.vertices = m.vertices.data();
            d.indices = m.indices.data();
            d.normals = m.has_normals ? m.normals.data() : nullptr;
            d.uvs = m.has_uvs ? m.uvs.data() : nullptr;
            d.tangents = m.has_tangents ? m.tangents.data() : nullptr;
            d.material_slots = m.material_slots.data();
            d.n_vertices = static_cast<uint32_t>(m.vertices.size() / 3);
            d.n_triangles = static_cast<uint32_t>(m.indices.size() / 3);
            d.n_material_slots = static_cast<uint32_t>(m.material_slots.size());
            d._pad = 0;
        }
        kinds.resize(kind_nodes.size());
        for (size_t i = 0; i < kind_nodes.size(); ++i) {
            kinds[i].nodes = kind_nodes[i].data();
            kinds[i].n_nodes = static_cast<uint32_t>(kind_nodes[i].size());
        }
        for (size_t i = 0; i < instances.size(); ++i) {
            instances[i].materials = instance_materials[i].data();
            instances[i].n_materials = static_cast<uint32_t>(instance_materials[i].size());
        }
        desc.abi_version = AKR_B200_ABI_VERSION;
        desc.n_meshes = static_cast<uint32_t>(meshes.size());
        desc.n_instances = static_cast<uint32_t>(instances.size());
        desc.n_shader_kinds = static_cast<uint32_t>(kinds.size());
        desc.meshes = meshes.data();
        desc.instances = instances.data();
        desc.shader_kinds = kinds.data();
        desc.shader_data = shader_data.data();
        desc.shader_data_size = shader_data.size();
        for (size_t i = 0; i < images.size(); ++i) images[i].texels = image_texels[i].data();
        desc.images = images.empty() ? nullptr : images.data();
        desc.n_images = static_cast<uint32_t>(images.size());
    }
};

namespace {

void load_scene_impl(const std::string &path, AkrHostScene &hs) {
    using akr::json::Value;
    const std::string dir = dirname_of(path);
    Value root = akr::json::parse(read_file(path));

    // ---- buffers (scene.rs:604-647) ----
    for (const auto &[name, b] : root.at("buffers").object()) {
        const std::string &type = b.at("type").string();
        std::vector<uint8_t> bytes;
        if (type == "path") {
            std::string p = b.at("path").string();
            std::string resolved;
            if (!p.empty() && p[0] == '/' && file_exists(p)) resolved = p;
            else if (file_exists(dir + "/" + p)) resolved = dir + "/" + p;
            else if (file_exists(dir + "/" + basename_any(p))) resolved = dir + "/" + basename_any(p);
            else throw std::runtime_error("buffer '" + name + "': cannot resolve path '" + p + "'");
            std::string s = read_file(resolved);
            bytes.assign(s.begin(), s.end());
            uint64_t expect = b.at("length").u64();
            if (expect != bytes.size())
                throw std::runtime_error("buffer size mismatch: expected " + std::to_string(expect) + ", got " +
                                         std::to_string(bytes.size()));
        } else if (type == "base64") {
            bytes = base64_decode(b.at("data").string());
        } else if (type == "binary") {
            for (const Value &v : b.at("data").array()) bytes.push_back(static_cast<uint8_t>(v.u32()));
        } else {
            throw std::runtime_error("buffer '" + name + "': unsupported type '" + type + "'");
        }
        hs.buffers[name] = std::move(bytes);
    }
    const Value &views = root.at("buffer_views");
    auto view_bytes = [&](const Value &ref, size_t elem, const char *what) -> std::pair<const uint8_t *, size_t> {
        const Value &v = views.at(ref.at("id").string());
        const std::vector<uint8_t> &buf = hs.buffers.at(v.at("buffer").at("id").string());
        size_t off = static_cast<size_t>(v.at("offset").u64());
        size_t len = static_cast<size_t>(v.at("length").u64());
        if (off > buf.size() || len > buf.size() - off) throw std::runtime_error(std::string("buffer view out of range for ") + what);  // (no off + len: it can wrap)
        if (len % elem != 0) throw std::runtime_error(std::string("Invalid slice length for ") + what);
        return {buf.data() + off, len};
    };

    // ---- geometries (load.rs:494-529); geom_id = BTreeMap order ----
    std::map<std::string, uint32_t> geom_ids;
    for (const auto &[name, g] : root.at("geometries").object()) {
        if (g.at("type").string() != "mesh") throw std::runtime_error("geometry '" + name + "': only meshes are supported");
        AkrHostScene::MeshStore m;
        auto copy_f = [&](const Value &ref, size_t elem, std::vector<float> &dst, const char *what) {
            auto [p, len] = view_bytes(ref, elem, what);
            dst.resize(len / 4);
            if (len) std::memcpy(dst.data(), p, len);
        };
        auto copy_u = [&](const Value &ref, size_t elem, std::vector<uint32_t> &dst, const char *what) {
            auto [p, len] = view_bytes(ref, elem, what);
            dst.resize(len / 4);
            if (len) std::memcpy(dst.data(), p, len);
        };
        copy_f(g.at("vertices"), 12, m.vertices, "vertices");
        copy_u(g.at("indices"), 12, m.indices, "indices");
        if (g.has("normals") && !g.at("normals").is_null()) {
            copy_f(g.at("normals"), 12, m.normals, "normals");
            m.has_normals = true;
        }
        if (g.has("uvs") && !g.at("uvs").is_null()) {
            copy_f(g.at("uvs"), 8, m.uvs, "uvs");
            m.has_uvs = true;
        }
        if (g.has("tangents") && !g.at("tangents").is_null()) {
            copy_f(g.at("tangents"), 12, m.tangents, "tangents");
            m.has_tangents = true;
        }
        copy_u(g.at("materials"), 4, m.material_slots, "materials");
        size_t ntri = m.indices.size() / 3;
        if (m.has_normals && m.normals.size() != ntri * 9) throw std::runtime_error("mesh '" + name + "': normals are not per-corner");
        if (m.has_uvs && m.uvs.size() != ntri * 6) throw std::runtime_error("mesh '" + name + "': uvs are not per-corner");
        if (m.has_tangents && m.tangents.size() != ntri * 9) throw std::runtime_error("mesh '" + name + "': tangents are not per-corner");
        geom_ids[name] = static_cast<uint32_t>(hs.mesh_store.size());
        hs.geometry_names.push_back(name);
        hs.mesh_store.push_back(std::move(m));
    }

    // ---- materials -> SVM (load.rs:242-253; compiler.rs:25-46) ----
    // image textures: one slot per distinct (buffer view, format, size, channels, extension, interpolation) (load.rs:611-646)
    std::map<std::string, uint32_t> image_slots;
    ImageResolver resolve_image = [&](const Value &image) -> uint32_t {
        const std::string view_id = image.at("data").at("id").string();
        const std::string &format = image.at("format").string();
        const std::string &extension = image.at("extension").string();
        const std::string &interp = image.at("interpolation").string();
        const uint32_t width = image.at("width").u32(), height = image.at("height").u32(), channels = image.at("channels").u32();
        const std::string key = view_id + "|" + format + "|" + extension + "|" + interp + "|" + std::to_string(width) + "x" + std::to_string(height) + "x" +
                                std::to_string(channels);
        auto it = image_slots.find(key);
        if (it != image_slots.end()) return it->second;
        AkrImage img{};
        if (extension == "repeat") img.address = AKR_ADDRESS_REPEAT;  // load.rs:684-689
        else if (extension == "clip") img.address = AKR_ADDRESS_ZERO;
        else if (extension == "mirror") img.address = AKR_ADDRESS_MIRROR;
        else if (extension == "extend") img.address = AKR_ADDRESS_EDGE;
        else throw std::runtime_error("image: unknown extension '" + extension + "'");
        // load.rs:690-699: Linear -> SamplerFilter::LinearLinear, Nearest -> SamplerFilter::LinearPoint, Cubic -> LinearLinear with
        // a warning.  In LuisaCompute's sampler LINEAR_POINT means "linear filtering inside a level, nearest level" (the D3D
        // MIN_MAG_LINEAR_MIP_POINT naming; only Filter::POINT fetches the nearest texel) and these textures have one level, so
        // ALL THREE modes sample bilinearly in the reference — reproduced here (third-party semantics, not vendored: ASSUMED).
        // AKR_FILTER_POINT stays in the ABI for hosts that bind a Filter::POINT sampler.
        if (interp == "nearest" || interp == "linear" || interp == "cubic") img.filter = AKR_FILTER_LINEAR;
        else throw std::runtime_error("image: unknown interpolation '" + interp + "'");
        auto [bytes, len] = view_bytes(image.at("data"), 1, "image data");
        std::vector<uint8_t> texels;
        decode_image(format, bytes, len, width, height, channels, img, texels);
        const uint32_t slot = static_cast<uint32_t>(hs.images.size());
        hs.images.push_back(img);
        hs.image_texels.push_back(std::move(texels));
        image_slots[key] = slot;
        return slot;
    };
    std::map<std::string, AkrShaderRef> mat_refs;
    for (const auto &[name, mat] : root.at("materials").object()) {
        CompiledShader cs = ShaderCompiler(mat.at("shader"), resolve_image).compile();
        uint32_t kind = UINT32_MAX;
        for (size_t k = 0; k < hs.kind_nodes.size(); ++k)
            if (same_bytecode(hs.kind_nodes[k], cs.nodes)) {
                kind = static_cast<uint32_t>(k);
                break;
            }
        if (kind == UINT32_MAX) {
            kind = static_cast<uint32_t>(hs.kind_nodes.size());
            hs.kind_nodes.push_back(cs.nodes);
        }
        AkrShaderRef r{kind, static_cast<uint32_t>(hs.shader_data.size())};
        hs.shader_data.insert(hs.shader_data.end(), cs.data.begin(), cs.data.end());
        size_t padding = 16 - (hs.shader_data.size() % 16);  // compiler.rs:38-41 (adds 16 when already aligned)
        hs.shader_data.insert(hs.shader_data.end(), padding, 0);
        mat_refs[name] = r;
        hs.material_names.push_back(name);
    }

    // ---- camera (load.rs:172-194) ----
    {
        const Value &cam = root.at("camera");
        if (cam.is_null()) throw std::runtime_error("scene has no camera");
        if (cam.at("type").string() != "perspective") throw std::runtime_error("only perspective cameras are supported");
        const Value &d = cam.at("data");
        Mat4 c2w = load_transform(d.at("transform"), true);
        c2w.store(hs.desc.camera.c2w);
        float fov_deg = d.at("fov").f32();
        hs.desc.camera.fov = fov_deg * (kPi / 180.0f);  // f32::to_radians
        float focal_distance = d.at("focal_distance").f32();
        float fstop = d.at("fstop").f32();
        hs.desc.camera.lens_radius = focal_distance / (2.0f * fstop);
        hs.desc.camera.focal_length = focal_distance;
        hs.desc.camera.width = d.at("sensor_width").u32();
        hs.desc.camera.height = d.at("sensor_height").u32();
        hs.desc.camera._pad = 0;
    }

    // ---- instances (load.rs:195-237,287-292) ----
    for (const auto &[name, inst] : root.at("instances").object()) {
        AkrInstance d{};
        const std::string &gname = inst.at("geometry").at("id").string();
        auto git = geom_ids.find(gname);
        if (git == geom_ids.end()) throw std::runtime_error("instance '" + name + "': unknown geometry '" + gname + "'");
        d.geom_id = git->second;
        load_transform(inst.at("transform"), false).store(d.transform);
        const AkrHostScene::MeshStore &m = hs.mesh_store[d.geom_id];
        uint32_t flags = 0;
        if (m.has_normals) flags |= AKR_MESH_HAS_NORMALS;
        if (m.has_uvs) flags |= AKR_MESH_HAS_UVS;
        if (m.has_tangents) flags |= AKR_MESH_HAS_TANGENTS;
        if (m.material_slots.size() > 1) flags |= AKR_MESH_HAS_MULTI_MATERIALS;
        d.flags = flags;
        std::vector<AkrShaderRef> mats;
        for (const Value &mr : inst.at("materials").array()) {
            auto mit = mat_refs.find(mr.at("id").string());
            if (mit == mat_refs.end()) throw std::runtime_error("instance '" + name + "': unknown material");
            mats.push_back(mit->second);
        }
        if (mats.empty()) throw std::runtime_error("instance '" + name + "' has no materials (mesh.rs:308)");
        hs.instance_materials.push_back(std::move(mats));
        hs.instances.push_back(d);
        hs.instance_names.push_back(name);
    }
    hs.refresh_desc();
}

void parse_task(const akr::json::Value &root_in, AkrRenderTask *t) {
    using akr::json::Value;
    akr_host_default_task(t);
    const Value *root = &root_in;
    if (root->is_array()) {  // RenderTask::Multi: this entry point handles the first config
        if (root->array().empty()) throw std::runtime_error("empty render task");
        root = &root->array()[0];
    }
    const Value &method = root->at("method");
    const std::string &type = method.at("type").string();
    if (type == "aov") {  // aov::Config (aov.rs:9-36)
        t->method = AKR_METHOD_AOV;
        if (method.has("spp")) t->aov.spp = method.at("spp").u32();
        if (method.has("remap")) t->aov.remap = method.at("remap").boolean() ? 1u : 0u;
        if (method.has("aov")) {
            static const char *names[] = {"ns", "ng", "tangent", "bitangent", "albedo", "roughness"};
            const std::string &a = method.at("aov").string();
            uint32_t k = 0;
            while (k < 6 && a != names[k]) ++k;
            if (k == 6) throw std::runtime_error("unknown aov '" + a + "'");
            t->aov.aov = k;
        }
    } else if (type != "pt") {
        throw std::runtime_error("method '" + type + "' is outside the hot-path scope (only 'pt' and 'aov')");
    }
    auto opt_u32 = [&](const char *k, uint32_t &dst) { if (method.has(k)) dst = method.at(k).u32(); };
    auto opt_bool = [&](const char *k, uint32_t &dst) { if (method.has(k)) dst = method.at(k).boolean() ? 1u : 0u; };
    opt_u32("spp", t->pt.spp);
    opt_u32("max_depth", t->pt.max_depth);
    opt_u32("spp_per_pass", t->pt.spp_per_pass);
    opt_u32("rr_depth", t->pt.rr_depth);
    opt_bool("use_nee", t->pt.use_nee);
    opt_bool("indirect_only", t->pt.indirect_only);
    opt_bool("force_diffuse", t->pt.force_diffuse);
    if (method.has("pixel_offset")) {
        t->pt.pixel_offset[0] = method.at("pixel_offset").at(0).i32();
        t->pt.pixel_offset[1] = method.at("pixel_offset").at(1).i32();
    }
    if (method.has("debug_depth") && !method.at("debug_depth").is_null()) t->pt.debug_depth = method.at("debug_depth").i32();
    if (root->has("sampler")) {
        const Value &s = root->at("sampler");
        const std::string &st = s.at("type").string();
        if (st == "pmj02bn") t->sampler.type = AKR_SAMPLER_PMJ02BN;
        else if (st == "independent") t->sampler.type = AKR_SAMPLER_INDEPENDENT;
        else throw std::runtime_error("unknown sampler '" + st + "'");
        t->sampler.seed = s.at("seed").u64();
    }
    const Value &film = root->at("film");
    if (film.has("out")) std::snprintf(t->out, sizeof(t->out), "%s", film.at("out").string().c_str());
    if (film.has("filter")) {
        const Value &f = film.at("filter");
        const std::string &ft = f.at("type").string();
        if (ft == "gaussian") t->filter.type = AKR_FILTER_GAUSSIAN;
        else if (ft == "box") t->filter.type = AKR_FILTER_BOX;
        else throw std::runtime_error("unknown filter '" + ft + "'");
        t->filter.radius = f.at("radius").f32();
    }
    if (film.has("color")) {
        const std::string &c = film.at("color").at("type").string();
        if (c != "srgb") throw std::runtime_error("film color '" + c + "' is todo!() in the reference (film.rs:190,227)");
    }
}

// ---- minimal OpenEXR writer: single-part scanline, NO_COMPRESSION, 3 x FLOAT channels ----------
void put_u32(std::vector<uint8_t> &o, uint32_t v) {
    for (int i = 0; i < 4; ++i) o.push_back(static_cast<uint8_t>((v >> (8 * i)) & 0xFF));
}
void put_u64(std::vector<uint8_t> &o, uint64_t v) {
    for (int i = 0; i < 8; ++i) o.push_back(static_cast<uint8_t>((v >> (8 * i)) & 0xFF));
}
void put_str(std::vector<uint8_t> &o, const char *s) {
    while (*s) o.push_back(static_cast<uint8_t>(*s++));
    o.push_back(0);
}
void put_f32(std::vector<uint8_t> &o, float f) {
    uint32_t v;
    std::memcpy(&v, &f, 4);
    put_u32(o, v);
}
void put_attr(std::vector<uint8_t> &o, const char *name, const char *type, const std::vector<uint8_t> &val) {
    put_str(o, name);
    put_str(o, type);
    put_u32(o, static_cast<uint32_t>(val.size()));
    o.insert(o.end(), val.begin(), val.end());
}

void write_exr(const std::string &path, const float *rgb, uint32_t w, uint32_t h) {
    std::vector<uint8_t> o;
    put_u32(o, 20000630u);  // magic
    put_u32(o, 2u);         // version 2, scanline, single part
    {
        std::vector<uint8_t> ch;
        for (const char *name : {"B", "G", "R"}) {  // channels sorted by name
            put_str(ch, name);
            put_u32(ch, 2u);  // FLOAT
            ch.push_back(0);  // pLinear
            ch.push_back(0);
            ch.push_back(0);
            ch.push_back(0);
            put_u32(ch, 1u);
            put_u32(ch, 1u);
        }
        ch.push_back(0);
        put_attr(o, "channels", "chlist", ch);
    }
    {
        std::vector<uint8_t> v{0};
        put_attr(o, "compression", "compression", v);
    }
    {
        std::vector<uint8_t> v;
        put_u32(v, 0);
        put_u32(v, 0);
        put_u32(v, w - 1);
        put_u32(v, h - 1);
        put_attr(o, "dataWindow", "box2i", v);
        put_attr(o, "displayWindow", "box2i", v);
    }
    {
        std::vector<uint8_t> v{0};
        put_attr(o, "lineOrder", "lineOrder", v);
    }
    {
        std::vector<uint8_t> v;
        put_f32(v, 1.0f);
        put_attr(o, "pixelAspectRatio", "float", v);
    }
    {
        std::vector<uint8_t> v;
        put_f32(v, 0.0f);
        put_f32(v, 0.0f);
        put_attr(o, "screenWindowCenter", "v2f", v);
    }
    {
        std::vector<uint8_t> v;
        put_f32(v, 1.0f);
        put_attr(o, "screenWindowWidth", "float", v);
    }
    o.push_back(0);  // end of header
    const uint64_t line_bytes = static_cast<uint64_t>(w) * 4u * 3u;
    const uint64_t table_pos = o.size();
    const uint64_t data_pos = table_pos + 8ull * h;
    for (uint32_t y = 0; y < h; ++y) put_u64(o, data_pos + y * (8ull + line_bytes));
    for (uint32_t y = 0; y < h; ++y) {
        put_u32(o, y);
        put_u32(o, static_cast<uint32_t>(line_bytes));
        for (int c : {2, 1, 0})  // B, G, R planes
            for (uint32_t x = 0; x < w; ++x) put_f32(o, rgb[(static_cast<size_t>(y) * w + x) * 3 + c]);
    }
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    f.write(reinterpret_cast<const char *>(o.data()), static_cast<std::streamsize>(o.size()));
}

void write_pfm(const std::string &path, const float *rgb, uint32_t w, uint32_t h) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    f << "PF\n" << w << " " << h << "\n-1.0\n";
    for (uint32_t y = 0; y < h; ++y)  // PFM rows are bottom-to-top
        f.write(reinterpret_cast<const char *>(rgb + static_cast<size_t>(h - 1 - y) * w * 3), static_cast<std::streamsize>(w) * 12);
}

}  // namespace

extern "C" {

int akr_host_load_scene(const char *scene_json_path, AkrHostScene **out_scene) {
    if (!scene_json_path || !out_scene) return set_error(AKR_ERR_INVALID_ARGUMENT, "null argument");
    *out_scene = nullptr;
    try {
        auto hs = std::make_unique<AkrHostScene>();
        load_scene_impl(scene_json_path, *hs);
        *out_scene = hs.release();
        g_last_error.clear();
        return AKR_OK;
    } catch (const std::exception &e) {
        return set_error(AKR_ERR_INVALID_ARGUMENT, std::string("akr_host_load_scene: ") + e.what());
    }
}

void akr_host_free_scene(AkrHostScene *scene) { delete scene; }

const AkrSceneDesc *akr_host_scene_desc(const AkrHostScene *scene) { return scene ? &scene->desc : nullptr; }

int akr_host_scene_set_resolution(AkrHostScene *scene, uint32_t width, uint32_t height) {
    if (!scene || width == 0 || height == 0) return set_error(AKR_ERR_INVALID_ARGUMENT, "bad resolution");
    scene->desc.camera.width = width;
    scene->desc.camera.height = height;
    return AKR_OK;
}

void akr_host_default_task(AkrRenderTask *t) {
    std::memset(t, 0, sizeof(*t));
    t->pt.spp = 256;
    t->pt.max_depth = 7;
    t->pt.rr_depth = 5;
    t->pt.spp_per_pass = 64;
    t->pt.use_nee = 1;
    t->pt.indirect_only = 0;
    t->pt.force_diffuse = 0;
    t->pt.pixel_offset[0] = t->pt.pixel_offset[1] = 0;
    t->pt.debug_depth = -1;
    t->sampler.type = AKR_SAMPLER_INDEPENDENT;  // SamplerConfig::default (sampler/mod.rs:291-295)
    t->sampler.seed = 0;
    t->filter.type = AKR_FILTER_GAUSSIAN;       // PixelFilter::default (film.rs:51-55)
    t->filter.radius = 1.5f;
    std::snprintf(t->out, sizeof(t->out), "out.exr");
    t->method = AKR_METHOD_PT;
    t->aov.spp = 256;  // aov::Config::default (aov.rs:29-36)
    t->aov.aov = AKR_AOV_SHADING_NORMAL;
    t->aov.remap = 1;
    t->aov._pad = 0;
}

int akr_host_parse_method_string(const char *method_json, AkrRenderTask *out_task) {
    if (!method_json || !out_task) return set_error(AKR_ERR_INVALID_ARGUMENT, "null argument");
    try {
        parse_task(akr::json::parse(method_json), out_task);
        g_last_error.clear();
        return AKR_OK;
    } catch (const std::exception &e) {
        std::string msg = e.what();
        int code = msg.find("outside the hot-path scope") != std::string::npos ? AKR_ERR_UNSUPPORTED : AKR_ERR_INVALID_ARGUMENT;
        return set_error(code, std::string("akr_host_parse_method: ") + msg);
    }
}

int akr_host_parse_method_file(const char *method_json_path, AkrRenderTask *out_task) {
    if (!method_json_path || !out_task) return set_error(AKR_ERR_INVALID_ARGUMENT, "null argument");
    try {
        std::string text = read_file(method_json_path);
        return akr_host_parse_method_string(text.c_str(), out_task);
    } catch (const std::exception &e) {
        return set_error(AKR_ERR_INVALID_ARGUMENT, std::string("akr_host_parse_method_file: ") + e.what());
    }
}

// util::write_image_ldr (util/mod.rs:64-94): linear -> sRGB (color.rs:565-571), (x * 255).clamp(0, 255) as u8 (truncation; NaN -> 0),
// 8-bit RGB.  Only the png container is written here (the reference hands the extension to image::RgbImage::save).
void write_png_ldr(const std::string &path, const float *rgb, uint32_t w, uint32_t h) {
    std::vector<uint8_t> raw((size_t)h * (1 + (size_t)w * 3));
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t *row = raw.data() + (size_t)y * (1 + (size_t)w * 3);
        row[0] = 0;  // filter type None
        for (uint32_t x = 0; x < w; ++x)
            for (int c = 0; c < 3; ++c) {
                const float l = rgb[((size_t)y * w + x) * 3 + c];
                const float s = l <= 0.0031308f ? l * 12.92f : std::pow(l, 1.0f / 2.4f) * 1.055f - 0.055f;
                float v = s * 255.0f;
                v = v != v ? 0.0f : (v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v));
                row[1 + (size_t)x * 3 + c] = (uint8_t)v;
            }
    }
    uLongf zlen = compressBound(static_cast<uLong>(raw.size()));
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), static_cast<uLong>(raw.size()), 6) != Z_OK) throw std::runtime_error("png: deflate failed");
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open '" + path + "' for writing");
    auto put32 = [](uint8_t *p, uint32_t v) {
        p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
    };
    auto chunk = [&](const char *type, const uint8_t *data, uint32_t n) {
        uint8_t hdr[8];
        put32(hdr, n);
        std::memcpy(hdr + 4, type, 4);
        f.write(reinterpret_cast<const char *>(hdr), 8);
        if (n) f.write(reinterpret_cast<const char *>(data), n);
        uLong crc = crc32(0L, reinterpret_cast<const Bytef *>(type), 4);
        if (n) crc = crc32(crc, data, n);
        uint8_t c[4];
        put32(c, (uint32_t)crc);
        f.write(reinterpret_cast<const char *>(c), 4);
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    f.write(reinterpret_cast<const char *>(sig), 8);
    uint8_t ihdr[13];
    put32(ihdr, w);
    put32(ihdr + 4, h);
    ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;  // 8 bit, RGB, deflate, adaptive filtering, no interlace
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", z.data(), (uint32_t)zlen);
    chunk("IEND", nullptr, 0);
}

int akr_host_write_image(const char *path, const float *rgb, uint32_t width, uint32_t height) {
    if (!path || !rgb || width == 0 || height == 0) return set_error(AKR_ERR_INVALID_ARGUMENT, "null argument");
    try {
        std::string p = path;
        if (p.size() >= 4 && p.substr(p.size() - 4) == ".exr") write_exr(p, rgb, width, height);
        else if (p.size() >= 4 && p.substr(p.size() - 4) == ".pfm") write_pfm(p, rgb, width, height);
        else if (p.size() >= 4 && p.substr(p.size() - 4) == ".png") write_png_ldr(p, rgb, width, height);
        else return set_error(AKR_ERR_UNSUPPORTED, "only .exr, .pfm and .png outputs are implemented");
        return AKR_OK;
    } catch (const std::exception &e) {
        return set_error(AKR_ERR_INVALID_ARGUMENT, std::string("akr_host_write_image: ") + e.what());
    }
}

const char *akr_host_last_error(void) { return g_last_error.c_str(); }

}  // extern "C"
