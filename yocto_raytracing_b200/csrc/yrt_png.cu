// yrt_png.cu — host-only: a parallel PNG encoder for the 8-bit frame.
//
// Replaces, as an option, save_image(filename, image4b) of the reference (src/image.cpp:41-44), which calls
// stb_image_write's single-threaded encoder (0.63 s for a 1920x1080 frame on this box — 45 frames' worth of rendering).
// PNG is lossless: any valid encoding of the same RGBA8 pixels decodes to the same image, so the parity contract
// (final PNG pixels) is untouched; the file bytes differ from stb's.
//
// Encoding: colour type 6 (RGBA), 8 bit, no interlace; every scanline is filtered with "Up" (type 2; "Sub" for the first
// line), which turns the smooth gradients of a rendered frame into runs of small values; the filtered lines are cut
// into bands, each band is deflated on its own host thread as an independent raw-deflate segment ending in a full flush
// (the last one with the final block), and the segments are concatenated behind one zlib header with the Adler-32 of
// the whole filtered stream combined from the bands' checksums — the construction pigz uses.  No GPU involved.
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "yrt_internal.h"

namespace yrt {
namespace {

void put_u32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}

void put_chunk(std::vector<uint8_t>& out, const char type[4], const uint8_t* data, size_t n) {
    put_u32(out, (uint32_t)n);
    size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    put_u32(out, (uint32_t)crc32(0L, out.data() + at, (uInt)(n + 4)));
}

struct Band {
    int row0 = 0, row1 = 0;
    std::vector<uint8_t> deflated;
    uLong adler = 1;
    size_t raw_len = 0;
    int status = Z_OK;
};

// filter rows [row0, row1) of the image and deflate them as one raw segment
void encode_band(const uint8_t* rgba, int width, int row0, int row1, bool last, int level, Band& b) {
    const size_t stride = (size_t)width * 4;
    std::vector<uint8_t> filt((size_t)(row1 - row0) * (stride + 1));
    for (int y = row0; y < row1; y++) {
        uint8_t* dst = filt.data() + (size_t)(y - row0) * (stride + 1);
        const uint8_t* cur = rgba + (size_t)y * stride;
        if (y == 0) {
            dst[0] = 1;   // Sub: byte minus the byte one pixel to the left
            for (size_t x = 0; x < stride; x++) dst[1 + x] = (uint8_t)(cur[x] - (x >= 4 ? cur[x - 4] : 0));
        } else {
            dst[0] = 2;   // Up: byte minus the byte above (raw pixels of the previous line, whichever band it belongs to)
            const uint8_t* up = cur - stride;
            for (size_t x = 0; x < stride; x++) dst[1 + x] = (uint8_t)(cur[x] - up[x]);
        }
    }
    b.raw_len = filt.size();
    b.adler = adler32(adler32(0L, Z_NULL, 0), filt.data(), (uInt)filt.size());
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    b.status = deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);   // raw deflate: header and trailer are written once, by the caller
    if (b.status != Z_OK) return;
    b.deflated.resize(deflateBound(&zs, (uLong)filt.size()) + 16);
    zs.next_in = filt.data(); zs.avail_in = (uInt)filt.size();
    zs.next_out = b.deflated.data(); zs.avail_out = (uInt)b.deflated.size();
    int rc = deflate(&zs, last ? Z_FINISH : Z_FULL_FLUSH);
    if ((last && rc != Z_STREAM_END) || (!last && (rc != Z_OK || zs.avail_in != 0))) b.status = rc == Z_OK ? Z_BUF_ERROR : rc;
    b.deflated.resize(b.deflated.size() - zs.avail_out);
    deflateEnd(&zs);
}

}  // namespace

int write_png_parallel(const char* path, const uint8_t* rgba, int width, int height, int threads, int level) {
    if (!path || !rgba || width <= 0 || height <= 0 || (size_t)width * 4 + 1 > 0x7fffffffu / 64) { set_error("yrt_write_png: bad arguments"); return YRT_ERR_INVALID; }
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const int min_rows = std::max(1, (int)((1u << 20) / ((size_t)width * 4)));       // at least ~1 MB per band: bands cost a flush each
    int n_bands = std::max(1, std::min(threads, (height + min_rows - 1) / min_rows));
    std::vector<Band> bands(n_bands);
    for (int k = 0; k < n_bands; k++) {
        bands[k].row0 = (int)((long long)height * k / n_bands);
        bands[k].row1 = (int)((long long)height * (k + 1) / n_bands);
    }
    std::vector<std::thread> pool;
    for (int k = 1; k < n_bands; k++)
        pool.emplace_back([&, k]() { encode_band(rgba, width, bands[k].row0, bands[k].row1, k == n_bands - 1, level, bands[k]); });
    encode_band(rgba, width, bands[0].row0, bands[0].row1, n_bands == 1, level, bands[0]);
    for (auto& t : pool) t.join();

    std::vector<uint8_t> idat;
    size_t total = 2 + 4;
    for (auto& b : bands) total += b.deflated.size();
    idat.reserve(total);
    idat.push_back(0x78); idat.push_back(0x01);          // zlib header: deflate, 32 KB window, no preset dictionary, check bits
    uLong adler = adler32(0L, Z_NULL, 0);
    for (auto& b : bands) {
        if (b.status != Z_OK) { set_error("yrt_write_png: deflate failed (%d)", b.status); return YRT_ERR_INVALID; }
        idat.insert(idat.end(), b.deflated.begin(), b.deflated.end());
        adler = adler32_combine(adler, b.adler, (z_off_t)b.raw_len);
    }
    put_u32(idat, (uint32_t)adler);

    std::vector<uint8_t> png;
    png.reserve(idat.size() + 64 + (idat.size() >> 20) * 12 + 64);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    png.insert(png.end(), sig, sig + 8);
    std::vector<uint8_t> ihdr;
    put_u32(ihdr, (uint32_t)width); put_u32(ihdr, (uint32_t)height);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8 bit, RGBA, deflate, adaptive filters, no interlace
    put_chunk(png, "IHDR", ihdr.data(), ihdr.size());
    const size_t piece = (size_t)1 << 30;                 // chunk length is a 31-bit field
    for (size_t off = 0; off < idat.size(); off += piece) put_chunk(png, "IDAT", idat.data() + off, std::min(piece, idat.size() - off));
    put_chunk(png, "IEND", nullptr, 0);

    FILE* f = fopen(path, "wb");
    if (!f) { set_error("yrt_write_png: cannot open %s", path); return YRT_ERR_INVALID; }
    size_t w = fwrite(png.data(), 1, png.size(), f);
    if (fclose(f) != 0 || w != png.size()) { set_error("yrt_write_png: short write to %s", path); return YRT_ERR_INVALID; }
    return YRT_OK;
}

}  // namespace yrt
