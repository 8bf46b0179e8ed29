// Host-side PNG decoding for the input stage: the reference reads every frame through PIL --
// `Image.open(path).convert('RGB')` for images (src/data_utils.py:134-165) and `np.array(Image.open(path))` for the 16-bit depth maps
// (src/data_utils.py:167-234) -- on DataLoader worker processes.  Here: PNG container parsing (CRC-checked), zlib inflate, the five
// scanline filters of the PNG specification (None / Sub / Up / Average / Paeth) and the colour-type conversion PIL's convert('RGB')
// performs, written straight into caller memory (e.g. the pinned staging buffers `ptta_input_stage` reads from).  No libpng, no PIL.
// Not supported (explicit error): Adam7 interlacing, bit depths 1/2/4, 16-bit colour.
#include <zlib.h>

#include <cstring>
#include <vector>

#include "common.cuh"
#include "../../include/ptta_b200.h"

namespace {

struct PngHeader { uint32_t w = 0, h = 0; int depth = 0, color = 0, interlace = 0, channels = 0; };

inline uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }

int channels_of(int color) {
    switch (color) { case 0: return 1; case 2: return 3; case 3: return 1; case 4: return 2; case 6: return 4; default: return 0; }
}

// walks the chunk list; fills the header, the palette and the concatenated IDAT stream
int parse(const unsigned char* f, size_t n, PngHeader& H, std::vector<unsigned char>* idat, unsigned char* palette /* 768 */, int* n_palette) {
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    PTTA_CHECK(f != nullptr && n >= 8 + 25 && memcmp(f, sig, 8) == 0, "png: not a PNG file (bad signature or %zu bytes)", n);
    size_t off = 8;
    bool have_ihdr = false, have_end = false;
    while (off + 12 <= n && !have_end) {
        const uint32_t len = be32(f + off);
        const unsigned char* type = f + off + 4;
        PTTA_CHECK((size_t)len <= n - off - 12, "png: chunk '%.4s' (%u bytes) runs past the end of the file", (const char*)type, len);
        const unsigned char* data = f + off + 8;
        const uint32_t crc = be32(data + len);
        const uint32_t want = (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, len + 4);
        PTTA_CHECK(crc == want, "png: CRC mismatch in chunk '%.4s'", (const char*)type);
        if (memcmp(type, "IHDR", 4) == 0) {
            PTTA_CHECK(len == 13 && !have_ihdr, "png: malformed IHDR");
            H.w = be32(data); H.h = be32(data + 4); H.depth = data[8]; H.color = data[9]; H.interlace = data[12];
            H.channels = channels_of(H.color);
            PTTA_CHECK(H.w > 0 && H.h > 0 && H.w <= (1u << 20) && H.h <= (1u << 20), "png: bad size %ux%u", H.w, H.h);
            PTTA_CHECK(data[10] == 0 && data[11] == 0, "png: unknown compression / filter method");
            PTTA_CHECK(H.channels != 0, "png: unknown colour type %d", H.color);
            have_ihdr = true;
        } else {
            PTTA_CHECK(have_ihdr, "png: chunk '%.4s' before IHDR", (const char*)type);
            if (memcmp(type, "PLTE", 4) == 0 && palette) {
                PTTA_CHECK(len % 3 == 0 && len <= 768, "png: malformed PLTE");
                memcpy(palette, data, len);
                *n_palette = (int)(len / 3);
            } else if (memcmp(type, "IDAT", 4) == 0 && idat) {
                idat->insert(idat->end(), data, data + len);
            } else if (memcmp(type, "IEND", 4) == 0) {
                have_end = true;
            }
        }
        off += 12 + (size_t)len;
    }
    PTTA_CHECK(have_ihdr, "png: no IHDR chunk");
    PTTA_CHECK(!idat || have_end, "png: truncated file (no IEND)");
    return 0;
}

inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// inflate + reverse the scanline filters; rows come out packed (no filter byte), `bpp` = bytes per complete pixel
int decode_rows(const std::vector<unsigned char>& idat, const PngHeader& H, std::vector<unsigned char>& rows) {
    PTTA_CHECK(H.interlace == 0, "png: Adam7 interlaced files are not supported");
    PTTA_CHECK(H.depth == 8 || H.depth == 16, "png: bit depth %d is not supported (8 or 16)", H.depth);
    const size_t bpp = (size_t)H.channels * (H.depth / 8), stride = (size_t)H.w * bpp;
    std::vector<unsigned char> raw((stride + 1) * H.h);
    z_stream zs; memset(&zs, 0, sizeof(zs));
    PTTA_CHECK(inflateInit(&zs) == Z_OK, "png: inflateInit failed");
    zs.next_in = const_cast<unsigned char*>(idat.data()); zs.avail_in = (uInt)idat.size();
    zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();
    const int rc = inflate(&zs, Z_FINISH);
    const size_t produced = raw.size() - zs.avail_out;
    inflateEnd(&zs);
    PTTA_CHECK(rc == Z_STREAM_END && produced == raw.size(), "png: corrupt image data (inflate returned %d after %zu of %zu bytes)", rc, produced, raw.size());
    rows.resize(stride * H.h);
    for (uint32_t y = 0; y < H.h; ++y) {
        const unsigned char* in = raw.data() + (stride + 1) * y;
        const int ft = in[0];
        ++in;
        unsigned char* cur = rows.data() + stride * y;
        const unsigned char* up = y ? cur - stride : nullptr;
        switch (ft) {
            case 0: memcpy(cur, in, stride); break;
            case 1:
                for (size_t i = 0; i < stride; ++i) cur[i] = (unsigned char)(in[i] + (i >= bpp ? cur[i - bpp] : 0));
                break;
            case 2:
                for (size_t i = 0; i < stride; ++i) cur[i] = (unsigned char)(in[i] + (up ? up[i] : 0));
                break;
            case 3:
                for (size_t i = 0; i < stride; ++i) cur[i] = (unsigned char)(in[i] + (((i >= bpp ? cur[i - bpp] : 0) + (up ? up[i] : 0)) >> 1));
                break;
            case 4:
                for (size_t i = 0; i < stride; ++i) {
                    const int a = i >= bpp ? cur[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
                    cur[i] = (unsigned char)(in[i] + paeth(a, b, c));
                }
                break;
            default: PTTA_CHECK(false, "png: unknown filter type %d in row %u", ft, y);
        }
    }
    return 0;
}

}  // namespace

extern "C" {

int ptta_png_info(const void* file_bytes, size_t n, int* width, int* height, int* channels, int* bit_depth) {
    PngHeader H;
    PTTA_TRY(parse((const unsigned char*)file_bytes, n, H, nullptr, nullptr, nullptr));
    if (width) *width = (int)H.w;
    if (height) *height = (int)H.h;
    if (channels) *channels = H.color == 3 ? 3 : H.channels;
    if (bit_depth) *bit_depth = H.depth;
    return 0;
}

// H x W x 3 bytes, what np.asarray(Image.open(path).convert('RGB')) holds: RGB as stored, RGBA / grey+alpha without the alpha channel, grey
// replicated, palette entries looked up
int ptta_png_decode_rgb8(const void* file_bytes, size_t n, unsigned char* out, size_t out_bytes) {
    PngHeader H; std::vector<unsigned char> idat, rows; unsigned char pal[768]; int npal = 0;
    memset(pal, 0, sizeof(pal));
    PTTA_TRY(parse((const unsigned char*)file_bytes, n, H, &idat, pal, &npal));
    PTTA_CHECK(H.depth == 8, "png: convert('RGB') of a %d-bit file is not supported", H.depth);
    PTTA_CHECK(out && out_bytes >= (size_t)H.w * H.h * 3, "png: output buffer of %zu bytes, %zu needed", out_bytes, (size_t)H.w * H.h * 3);
    PTTA_CHECK(H.color != 3 || npal > 0, "png: palette image without PLTE");
    PTTA_TRY(decode_rows(idat, H, rows));
    const size_t px = (size_t)H.w * H.h;
    const unsigned char* r = rows.data();
    switch (H.color) {
        case 2: memcpy(out, r, px * 3); break;
        case 6: for (size_t i = 0; i < px; ++i) { out[3 * i] = r[4 * i]; out[3 * i + 1] = r[4 * i + 1]; out[3 * i + 2] = r[4 * i + 2]; } break;
        case 0: for (size_t i = 0; i < px; ++i) out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = r[i]; break;
        case 4: for (size_t i = 0; i < px; ++i) out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = r[2 * i]; break;
        case 3:
            for (size_t i = 0; i < px; ++i) {
                PTTA_CHECK(r[i] < npal, "png: palette index %d out of range (%d entries)", r[i], npal);
                memcpy(out + 3 * i, pal + 3 * r[i], 3);
            }
            break;
    }
    return 0;
}

// H x W host-endian uint16, what np.array(Image.open(path)) holds for a grey file ('I;16' for 16-bit depth maps, 'L' for 8-bit ones)
int ptta_png_decode_gray16(const void* file_bytes, size_t n, unsigned short* out, size_t out_bytes) {
    PngHeader H; std::vector<unsigned char> idat, rows;
    PTTA_TRY(parse((const unsigned char*)file_bytes, n, H, &idat, nullptr, nullptr));
    PTTA_CHECK(H.color == 0, "png: depth maps are single-channel grey files, this one has colour type %d", H.color);
    PTTA_CHECK(out && out_bytes >= (size_t)H.w * H.h * 2, "png: output buffer of %zu bytes, %zu needed", out_bytes, (size_t)H.w * H.h * 2);
    PTTA_TRY(decode_rows(idat, H, rows));
    const size_t px = (size_t)H.w * H.h;
    const unsigned char* r = rows.data();
    if (H.depth == 16) for (size_t i = 0; i < px; ++i) out[i] = (unsigned short)((r[2 * i] << 8) | r[2 * i + 1]);
    else for (size_t i = 0; i < px; ++i) out[i] = r[i];
    return 0;
}

}  // extern "C"
